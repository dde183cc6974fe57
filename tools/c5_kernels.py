"""Kernel census of one C5 training step (torch.profiler; development tool)."""
import collections, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import profile, ProfilerActivity
import graddft_b200 as gd
from graddft_b200.synthetic import synthetic_molecule
import bench
dev = torch.device("cuda:0")
shapes = bench._train_shapes()
mols = [gd.molecule_from_tensors(synthetic_molecule(N, n, n_omega=2, seed=1993 + i, device=dev, mask_frac=0.0), dev) for i, (N, n) in enumerate(shapes)]
for m in mols: m.packed_basis
fun = gd.DM21()
params = {k: v.requires_grad_(True) for k, v in fun.generate_DM21_weights(device=dev).items()}
leaves = list(params.values())
predictor = gd.non_scf_predictor(fun)
def step():
    energies = predictor.energy_only_batch(params, mols)
    total = sum(((e + 1.0) / m.mo_occ.sum()) ** 2 for e, m in zip(energies, mols))
    return torch.autograd.grad(total / 64, leaves)
for _ in range(2): step()
torch.cuda.synchronize()
t0 = time.perf_counter(); step(); t_enq = time.perf_counter() - t0; torch.cuda.synchronize(); t_all = time.perf_counter() - t0
print(f"step wall {t_all * 1e3:.1f} ms, host enqueue {t_enq * 1e3:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
c = collections.Counter(); t = collections.Counter()
for ev in prof.events():
    if "cuda" in str(ev.device_type).lower():
        c[ev.name] += 1; t[ev.name] += ev.device_time
tot = sum(t.values())
print(f"{sum(c.values())} kernels, {tot / 1e3:.1f} ms of device time")
for k, us in t.most_common(22):
    print(f"{us / 1e3:8.2f} ms {100 * us / tot:5.1f}%  x{c[k]:4d}  {k[:120]}")
