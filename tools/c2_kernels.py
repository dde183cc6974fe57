"""Kernel census of ONE SCF iteration (difference between a 3-cycle and a 2-cycle eager loop) with torch.profiler, and the
in-graph iteration time (development tool).  python tools/c2_kernels.py [--shape c2]"""
import argparse, collections, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import profile, ProfilerActivity
import graddft_b200 as gd
import bench

ap = argparse.ArgumentParser(); ap.add_argument("--shape", default="c2"); a = ap.parse_args()
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES[a.shape]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev)

def census(cycles):
    loop = gd.diff_scf_loop(gd.B3LYP, cycles=cycles)
    with torch.no_grad():
        loop(None, m); torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            loop(None, m); torch.cuda.synchronize()
    c = collections.Counter(); t = collections.Counter()
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower():
            c[ev.name] += 1; t[ev.name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
    return c, t

c6, t6 = census(6); c5, t5 = census(5)
rows = []
for k in c6:
    dn = c6[k] - c5.get(k, 0)
    if dn: rows.append((t6[k] - t5.get(k, 0), dn, k))
rows.sort(reverse=True)
print(f"one iteration (cycle 6): {sum(r[1] for r in rows)} kernels/memcpys, {sum(r[0] for r in rows):.1f} us of device time")
for us, dn, k in rows:
    print(f"{us:8.1f} us  x{dn:2d}  {k[:150]}")
