"""Where one SCF iteration goes (development tool): per-C-ABI-call CUDA-event times inside one predict, plus the
n x n host-framework pieces, at a benzene-shaped or H2O-shaped synthetic molecule."""
import sys, time, argparse
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
from graddft_b200 import ops
sys.argv = [sys.argv[0]] + sys.argv[1:]
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="c3")
a = ap.parse_args()
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES[a.shape]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev)
pred = gd.energy_predictor(gd.B3LYP)
for _ in range(2): e, f = pred(None, m)
torch.cuda.synchronize()
ops.TIMING = {}
t0 = time.perf_counter(); e, f = pred(None, m); torch.cuda.synchronize(); t1 = time.perf_counter()
tm, ops.TIMING = ops.TIMING, None
print(f"predict wall {1e3*(t1-t0):.3f} ms")
tot = 0.0
for k, ev in tm.items():
    ms = [x.elapsed_time(y) for x, y in ev]; tot += sum(ms)
    print(f"  {k:22s} calls {len(ms)}  " + " ".join(f"{v:.3f}" for v in ms))
print(f"  timed kernels total {tot:.3f} ms")
def wall(fn, rep=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep * 1e3
print(f"safe_fock_solver {wall(lambda: gd.safe_fock_solver(f, m.s1e)):.3f} ms")
print(f"eigh only        {wall(lambda: torch.linalg.eigh(f)):.3f} ms")
print(f"cholesky+inv     {wall(lambda: torch.linalg.inv(torch.linalg.cholesky(m.s1e))):.3f} ms")
n = sh["n"]
A = torch.eye(n, dtype=torch.float64, device=dev)
diis = gd.evaluate.JittableDiis(m.s1e, A, 10)
z = torch.zeros((10, 2, n, n), dtype=torch.float64, device=dev)
data = (z, z.clone(), torch.zeros(10, dtype=torch.float64, device=dev), z.clone())
print(f"diis.run         {wall(lambda: diis.run((m.rdm1, f, e), data, 0)):.3f} ms")
mm = m.replace(fock=f)
print(f"get_occ+make_rdm1+mo_grads {wall(lambda: (mm.get_occ(), mm.make_rdm1(), torch.linalg.norm(mm.get_mo_grads()))):.3f} ms")
