"""Kernel census of one DM21 energy_predictor call at the benzene shape (torch.profiler; development tool)."""
import collections, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import profile, ProfilerActivity
import graddft_b200 as gd
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c3_dm21"]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev, n_omega=2)
fun = gd.DM21(); params = fun.generate_DM21_weights(device=dev); pred = gd.energy_predictor(fun)
with torch.no_grad():
    for _ in range(3): pred(params, m)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        pred(params, m); torch.cuda.synchronize()
c = collections.Counter(); t = collections.Counter()
for ev in prof.events():
    if "cuda" in str(ev.device_type).lower():
        c[ev.name] += 1; t[ev.name] += ev.device_time
tot = sum(t.values())
print(f"{sum(c.values())} kernels, {tot / 1e3:.2f} ms of device time")
for k, us in t.most_common(30):
    print(f"{us / 1e3:8.3f} ms {100 * us / tot:5.1f}%  x{c[k]:3d}  {k[:130]}")
