"""SCF-iteration latency probe (development tool): H2O- and benzene-shaped synthetic molecules."""
import sys, time, argparse
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
from graddft_b200.synthetic import synthetic_molecule

def sync_time(fn, warm=2, rep=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep * 1e3

ap = argparse.ArgumentParser(); ap.add_argument("--N", type=int, default=34000); ap.add_argument("--n", type=int, default=43)
ap.add_argument("--W", type=int, default=1); ap.add_argument("--func", default="B3LYP"); ap.add_argument("--cycles", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda:0")
mol = synthetic_molecule(a.N, a.n, n_omega=a.W, seed=1984, device=dev, mask_frac=0.0)
if a.W: mol["omegas"] = [0.0, 0.4][:a.W]
m = gd.molecule_from_tensors(mol, dev)
m.packed_basis
f = getattr(gd, a.func) if a.func != "DM21" else gd.DM21()
params = f.generate_DM21_weights(device=dev) if a.func == "DM21" else None
pred = gd.energy_predictor(f)
print("predict        ms:", sync_time(lambda: pred(params, m)))
print("xc build       ms:", sync_time(lambda: gd.xc_energy_and_grads(f, params, m.rdm1, m, create_graph=False)))
print("eigh           ms:", sync_time(lambda: gd.safe_fock_solver(m.rdm1, m.s1e)))
P = m.rdm1.sum(0)
print("J sweep        ms:", sync_time(lambda: gd.ops.coulomb_j_and_energy(P, m.rep_tensor)))
loop = gd.diff_scf_loop(f, cycles=a.cycles)
t = sync_time(lambda: loop(params, m), warm=1, rep=3)
print(f"diff_scf_loop({a.cycles}) ms: {t:.2f}  -> {t/ (a.cycles+1):.3f} ms/iter, {1e3*(a.cycles+1)/t:.1f} iter/s")
