"""One eager 3-cycle SCF loop at the benzene shape with the profiler range around it (run under ncu -k regex:... to capture
the ERI sweep and the cluster eigensolver at their real shapes)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c3"]
m = bench._scf_shard(int(sys.argv[1]) if len(sys.argv) > 1 else 60000, sh["n"], 0, 1, dev)
loop = gd.diff_scf_loop(gd.B3LYP, cycles=3)
with torch.no_grad():
    loop(None, m); torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loop(None, m); torch.cuda.synchronize()
    torch.cuda.profiler.stop()
