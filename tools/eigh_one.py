"""Development probe: gdft_sym_eigh at n = 43 (for ncu)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 43
g = torch.Generator().manual_seed(n)
A = torch.randn(2, n, n, generator=g, dtype=torch.float64)
A = (A + A.transpose(1, 2)).to(dev)
for _ in range(3):
    w, V = ops.sym_eigh(A)
torch.cuda.synchronize()
