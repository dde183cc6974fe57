"""K1 row-tile shape probe: 128-row vs 64-row CTAs at grid sizes around the wave-quantisation regime (development tool)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
from graddft_b200._lib import GDFT_RHO, GDFT_GRAD, GDFT_TAU, GDFT_LAPL, GDFT_HF
from graddft_b200.synthetic import synthetic_molecule

def timeit(fn, warm=2, rep=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]

dev = torch.device("cuda:0")
for (N, n) in [(62_500, 264), (125_000, 264), (250_000, 264), (500_000, 264), (34_000, 43), (250_000, 400), (40_000, 100), (100_000, 43)]:
    mol = synthetic_molecule(N, n, n_omega=1, seed=1984, device=dev, with_eri=False, with_grad2=True)
    basis = ops.PackedBasis(mol["ao"], mol["grad_ao"], mol["grad_n_ao2"], mol["chi"]); D = mol["rdm1"]; del mol
    unit = 2.0 * N * n * n
    line = f"N={N:7d} n={n:3d}"
    for flags, units, name in ((GDFT_RHO | GDFT_GRAD, 2, "GGA"), (GDFT_RHO | GDFT_GRAD | GDFT_LAPL | GDFT_HF, 8, "B3LYP")):
        for rows in ("128", "64", "auto"):
            if rows == "auto": os.environ.pop("GDFT_FWD_ROWS", None)
            else: os.environ["GDFT_FWD_ROWS"] = rows
            ms = timeit(lambda: ops._density_fwd_raw(basis, D, flags))
            line += f" | {name} rows={rows:4s} {ms:8.3f} ms {units*unit/ms/1e9:6.2f} TF"
    print(line, flush=True)
    del basis; torch.cuda.empty_cache()
