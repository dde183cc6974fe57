"""Development probe: one gdft_chi_contract launch at the bench shape (for ncu)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
dev = torch.device("cuda:0")
n, rows = 264, 9472
g = torch.Generator(device=dev).manual_seed(1)
ao = torch.randn(rows, n, generator=g, dtype=torch.float64, device=dev)
D = torch.randn(2, n, n, generator=g, dtype=torch.float64, device=dev)
nu = torch.randn(rows, n, n, generator=g, dtype=torch.float64, device=dev)
chi = torch.empty((rows, 1, 2, n), dtype=torch.float64, device=dev)
for _ in range(3):
    ops.chi_contract_(chi, 0, 0, ao, D, nu)
torch.cuda.synchronize()
