"""One DM21 energy_predictor call at the benzene shape inside a profiler range (for ncu -k regex:density_bwd: the XC VJP with
tau terms and the omega-summed exact-exchange GEMM)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import graddft_b200 as gd
import bench
dev = torch.device("cuda:0")
sh = bench.SCF_SHAPES["c3_dm21"]
m = bench._scf_shard(sh["N"], sh["n"], 0, 1, dev, n_omega=2)
fun = gd.DM21(); params = fun.generate_DM21_weights(device=dev); pred = gd.energy_predictor(fun)
with torch.no_grad():
    for _ in range(2): pred(params, m)
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    pred(params, m); torch.cuda.synchronize(); torch.cuda.profiler.stop()
