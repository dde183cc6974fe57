"""One J+K sweep and one K transpose at n = 264 (38.9 GB tensor) inside a profiler range, for
`ncu --set full --profile-from-start off -k regex:"eri_jk_kernel|eri_kt_kernel"`.  python tools/eri_jk_ncu.py [n]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 264
g = torch.Generator(device=dev).manual_seed(n)
eri = torch.rand((n, n, n, n), generator=g, dtype=torch.float64, device=dev)
P = torch.randn((n, n), generator=g, dtype=torch.float64, device=dev)
with torch.no_grad():
    ops._eri_jk_raw(P, eri); ops._eri_kt_raw(P, eri)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    ops._eri_jk_raw(P, eri); ops._eri_kt_raw(P, eri)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
print(f"n={n}: algorithmic bytes per sweep {8 * n ** 4 / 1e9:.3f} GB")
