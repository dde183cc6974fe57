import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
dev = torch.device("cuda:0")
for n in (12, 43):
    g = torch.Generator().manual_seed(n)
    A = torch.randn(2, n, n, generator=g, dtype=torch.float64); A = (A + A.transpose(1, 2)).to(dev)
    for dbg in (0, 8, 9, 10, 12, 15):
        os.environ["GDFT_EIG_DBG"] = str(dbg)
        for _ in range(3): ops.sym_eigh(A)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ops.sym_eigh(A)
        e1.record(); torch.cuda.synchronize()
        print(f"n={n} dbg={dbg:2d} (8=fixed 8 sweeps, +1 no V, +2 no rotation math, +4 no A write): {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us")
