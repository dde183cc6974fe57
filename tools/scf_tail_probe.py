"""Per-cycle anatomy of the SCF loop's n x n tail (development tool): sweeps and time of the eigensolver in every cycle,
time of the Fock build, and what is left (DIIS, occupations, rdm1, orbital gradient).
    python tools/scf_tail_probe.py --shape c3 --rows 20000 --cycles 22"""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
from graddft_b200 import ops, evaluate
import bench

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="c3"); ap.add_argument("--rows", type=int, default=20000); ap.add_argument("--cycles", type=int, default=22)
a = ap.parse_args()
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES[a.shape]
m = bench._scf_shard(min(a.rows, sh["N"]), sh["n"], 0, 1, dev)
loop = gd.diff_scf_loop(gd.B3LYP, cycles=a.cycles)
infos = []
orig = ops.sym_eigh
def spy(C, V0=None, info=None):
    info = torch.zeros(C.shape[0], dtype=torch.int32, device=C.device)
    infos.append(info)
    return orig(C, V0, info)
ops.sym_eigh = spy
with torch.no_grad():
    loop(None, m); torch.cuda.synchronize()
    infos.clear()
    ops.TIMING = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = loop(None, m); e1.record(); torch.cuda.synchronize()
    tm, ops.TIMING = ops.TIMING, None
tot = e0.elapsed_time(e1)
per = {k: [x.elapsed_time(y) for x, y in ev] for k, ev in tm.items()}
print(f"loop of {a.cycles} cycles: {tot:.3f} ms eager, E = {float(out.energy):.10f}")
for k, v in per.items():
    print(f"  {k:24s} calls {len(v):3d} total {sum(v):8.3f} ms  mean {sum(v)/len(v):.4f}")
eig = per.get("gdft_sym_eigh", [])
print("sweeps per cycle:", [i.tolist() for i in infos])
print("eigh ms per cycle:", [round(x, 3) for x in eig])
timed = sum(sum(v) for v in per.values())
print(f"timed library calls {timed:.3f} ms; rest (host-framework n x n glue + gaps) {tot - timed:.3f} ms = {(tot - timed) / a.cycles:.4f} ms per cycle")
