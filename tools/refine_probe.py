"""Development probe: per-cycle eigensolve path and time in the benzene-shaped SCF loop (refinement accepted / refused)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
from graddft_b200 import evaluate
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c3"]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev)
orig = evaluate.refine_eigh
log = []
def spy(C, X):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = orig(C, X)
    torch.cuda.synchronize(); log.append(("accepted" if r[0] is not None else "refused", (time.perf_counter() - t0) * 1e3))
    return r
evaluate.refine_eigh = spy
for cycles in (2, 6, 10):
    log.clear()
    loop = gd.diff_scf_loop(gd.B3LYP, cycles=cycles)
    with torch.no_grad():
        loop(None, m); torch.cuda.synchronize(); log.clear()
        t0 = time.perf_counter(); out = loop(None, m); torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"cycles={cycles}: {1e3 * (t1 - t0):.1f} ms total, E={float(out.energy):.10f}", [f"{a}:{t:.2f}ms" for a, t in log])
