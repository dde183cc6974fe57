"""Development probe: eager vs CUDA-graph-replayed SCF loop at the H2O shape."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c2"]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev)
def wall(fn, rep=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep * 1e3
for cycles in (2, 6):
    eager = gd.diff_scf_loop(gd.B3LYP, cycles=cycles)
    jit = gd.make_jitted_scf_loop(gd.B3LYP, cycles=cycles)
    with torch.no_grad():
        e0 = eager(None, m); e_eager = float(e0.energy); r_eager = e0.rdm1.clone()
        e1 = jit(None, m); torch.cuda.synchronize()
        print(f"cycles={cycles} eager E={e_eager:.12f} graph E={float(e1.energy):.12f} max|drdm1|={float((e1.rdm1 - r_eager).abs().max()):.2e}")
        print(f"   eager {wall(lambda: eager(None, m)):.3f} ms   graph {wall(lambda: jit(None, m)):.3f} ms")
