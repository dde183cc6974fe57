"""Development probe: the kernels of one graph-replayed H2O-shaped SCF loop (run under ncu for the launch list)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c2"]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev)
cycles = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eager = gd.diff_scf_loop(gd.B3LYP, cycles=cycles)
with torch.no_grad():
    for _ in range(2):
        eager(None, m)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eager(None, m)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
