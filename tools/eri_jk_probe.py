"""CUDA-event timings of the ERI sweeps: J only, J+K from one pass (gdft_eri_jk with K), the K transpose -- against the bytes of
the tensor (8 n^4).  python tools/eri_jk_probe.py [n ...]   (default 43 128 264)"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops

dev = torch.device("cuda:0")


def timed(fn, reps):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n in [int(x) for x in sys.argv[1:]] or [43, 128, 264]:
    g = torch.Generator(device=dev).manual_seed(n)
    eri = torch.rand((n, n, n, n), generator=g, dtype=torch.float64, device=dev)
    P = torch.randn((n, n), generator=g, dtype=torch.float64, device=dev)
    gb = 8 * n ** 4 / 1e9
    reps = 20 if gb < 2 else 5
    with torch.no_grad():
        J, K = ops._eri_jk_raw(P, eri)
        J1 = ops._eri_j_raw(P, eri)[0]
        rows = torch.randint(0, n, (4,), generator=torch.Generator().manual_seed(1)).tolist()
        for p in rows:  # spot check against einsum on a few p
            Kp = torch.einsum("qrt,qt->r", eri[p], P)
            assert float((K[p] - Kp).abs().max() / Kp.abs().max()) < 1e-12
        dj = float((J - J1).abs().max() / J1.abs().max())
        t_j = timed(lambda: ops._eri_j_raw(P, eri), reps)
        t_jk = timed(lambda: ops._eri_jk_raw(P, eri), reps)
        t_kt = timed(lambda: ops._eri_kt_raw(P, eri), reps)
    print(f"n={n}: tensor {gb:.3f} GB | J only {t_j:.3f} ms = {gb / t_j:.2f} TB/s | J+K one pass {t_jk:.3f} ms = {gb / t_jk:.2f} TB/s | "
          f"K transpose {t_kt:.3f} ms = {gb / t_kt:.2f} TB/s | max rel |J(J+K) - J(J only)| = {dj:.1e}")
    del eri
    torch.cuda.empty_cache()
