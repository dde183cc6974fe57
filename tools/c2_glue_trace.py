"""Which Python lines launch the host-framework (ATen) kernels of one B3LYP predictor call at the H2O shape (development tool)."""
import collections, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import profile, ProfilerActivity
import graddft_b200 as gd
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c2"]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev)
pred = gd.energy_predictor(gd.B3LYP)
with torch.no_grad():
    for _ in range(3): pred(None, m)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        pred(None, m); torch.cuda.synchronize()
rows = collections.Counter()
for ev in prof.events():
    if ev.name.startswith("aten::") and ev.device_time_total > 0 and not ev.cpu_children:
        frames = [f for f in (ev.stack or []) if "graddft_b200" in f or "bench.py" in f]
        rows[(ev.name, frames[0].strip() if frames else "?")] += 1
for (name, where), cnt in sorted(rows.items(), key=lambda kv: kv[0][1]):
    print(f"x{cnt}  {name:28s} {where[:150]}")
