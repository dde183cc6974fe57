import sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, argparse
import graddft_b200 as gd
from graddft_b200.synthetic import synthetic_molecule
from torch.profiler import profile, ProfilerActivity
ap = argparse.ArgumentParser(); ap.add_argument("--N", type=int, default=500000); ap.add_argument("--n", type=int, default=264)
ap.add_argument("--W", type=int, default=1); ap.add_argument("--func", default="B3LYP")
a = ap.parse_args()
dev = torch.device("cuda:0")
mol = synthetic_molecule(a.N, a.n, n_omega=a.W, seed=1984, device=dev, mask_frac=0.0, with_eri=(a.n <= 264))
if a.W: mol["omegas"] = [0.0, 0.4][:a.W]
m = gd.molecule_from_tensors(mol, dev); m.packed_basis
f = getattr(gd, a.func) if a.func != "DM21" else gd.DM21()
params = f.generate_DM21_weights(device=dev) if a.func == "DM21" else None
pred = gd.energy_predictor(f)
for _ in range(2): pred(params, m)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): pred(params, m)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
