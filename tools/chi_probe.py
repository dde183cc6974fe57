"""Development probe: gdft_chi_contract GB/s per occupancy variant (GDFT_CHI_PER_SM) and shape."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
dev = torch.device("cuda:0")
for n, rows in ((264, 9472), (264, 8192), (264, 1024), (400, 4736), (128, 18944), (44, 37888), (512, 2368)):
    g = torch.Generator(device=dev).manual_seed(1)
    ao = torch.randn(rows, n, generator=g, dtype=torch.float64, device=dev)
    D = torch.randn(2, n, n, generator=g, dtype=torch.float64, device=dev)
    nu = torch.randn(rows, n, n, generator=g, dtype=torch.float64, device=dev)
    chi = torch.empty((rows, 1, 2, n), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for v in ("0", "1"):
        os.environ["GDFT_CHI_TMA"] = v
        for _ in range(2):
            ops.chi_contract_(chi, 0, 0, ao, D, nu)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.chi_contract_(chi, 0, 0, ao, D, nu); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"n={n} rows={rows} tma={v}: {ms:.3f} ms  {8.0 * rows * n * n / ms / 1e6:.0f} GB/s")
    del nu
