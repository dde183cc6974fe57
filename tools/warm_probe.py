"""Development probe: how diagonal is V_{k-1}^T C_k V_{k-1} in the H2O-shaped SCF loop (would a warm-started Jacobi pay?)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
from graddft_b200 import evaluate
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c2"]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev)
log = []
orig = evaluate.safe_eigh
prev = {}
def spy(C):
    w, V = orig(C)
    if "V" in prev:
        Ap = prev["V"].transpose(1, 2) @ C @ prev["V"]
        off = (Ap ** 2).sum((1, 2)) - (torch.diagonal(Ap, dim1=1, dim2=2) ** 2).sum(1)
        log.append((off / (Ap ** 2).sum((1, 2))).tolist())
    prev["V"] = V
    return w, V
evaluate.safe_eigh = spy
with torch.no_grad():
    gd.diff_scf_loop(gd.B3LYP, cycles=12)(None, m)
for k, r in enumerate(log):
    print(f"cycle {k + 1}: off^2/||A||^2 after rotating by the previous eigenvectors = {r[0]:.2e}, {r[1]:.2e}")
