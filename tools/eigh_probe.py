"""Development probe: gdft_sym_eigh timing / sweeps at the SCF shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
dev = torch.device("cuda:0")
for n in (2, 12, 43, 64, 72, 90):
    g = torch.Generator().manual_seed(n)
    A = torch.randn(2, n, n, generator=g, dtype=torch.float64)
    A = (A + A.transpose(1, 2)).to(dev)
    for _ in range(3):
        w, V = ops.sym_eigh(A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        w, V = ops.sym_eigh(A)
    e1.record(); torch.cuda.synchronize()
    wr, Vr = torch.linalg.eigh(A)
    torch.cuda.synchronize()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(5):
        torch.linalg.eigh(A)
    c1.record(); torch.cuda.synchronize()
    lib_us = c0.elapsed_time(c1) / 5 * 1e3
    res = (A @ V - V * w.unsqueeze(-2)).abs().max().item()
    print(f"n={n:4d}  {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us/call  max|w-w_ref|={float((w - wr).abs().max()):.2e}  resid={res:.2e}  (torch.linalg.eigh {lib_us:.0f} us)")
