"""Tile-shape tuning probe: times density fwd/bwd for forced tile sizes (GDFT_FWD_NTS / GDFT_BWD_MT)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
from graddft_b200._lib import GDFT_RHO, GDFT_GRAD, GDFT_TAU, GDFT_LAPL
from graddft_b200.synthetic import synthetic_molecule

def timeit(fn, warm=2, rep=4):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]

dev = torch.device("cuda:0")
for (N, n) in [(500_000, 264), (500_000, 240), (300_000, 384), (1_000_000, 43), (500_000, 100)]:
    mol = synthetic_molecule(N, n, seed=1984, device=dev, with_eri=False, with_grad2=True)
    basis = ops.PackedBasis(mol["ao"], mol["grad_ao"], mol["grad_n_ao2"]); D = mol["rdm1"]; del mol
    unit = 2.0 * N * n * n
    rb = torch.randn(N, 2, dtype=torch.float64, device=dev); gb = torch.randn(N, 2, 3, dtype=torch.float64, device=dev)
    lb = torch.randn(N, 2, dtype=torch.float64, device=dev)
    for nts in (1, 2, 3, 4, 5):
        os.environ["GDFT_FWD_NTS"] = str(nts)
        ms = timeit(lambda: ops._density_fwd_raw(basis, D, GDFT_RHO | GDFT_GRAD))
        ms2 = timeit(lambda: ops._density_fwd_raw(basis, D, GDFT_RHO | GDFT_GRAD | GDFT_LAPL))
        print(f"n={n} fwd NTS={nts}: GGA {ms:8.3f} ms {2*unit/ms/1e9:6.2f} TF | +LAPL {ms2:8.3f} ms {8*unit/ms2/1e9:6.2f} TF", flush=True)
    del os.environ["GDFT_FWD_NTS"]
    for mt in (0,):
        pass
        ms = timeit(lambda: ops._density_bwd_raw(basis, GDFT_RHO | GDFT_GRAD, rb, gb, None, None))
        ms2 = timeit(lambda: ops._density_bwd_raw(basis, GDFT_RHO | GDFT_GRAD | GDFT_LAPL, rb, gb, None, lb))
        print(f"n={n} bwd MT={mt}: GGA {ms:8.3f} ms {2*unit/ms/1e9:6.2f} TF | +LAPL {ms2:8.3f} ms {8*unit/ms2/1e9:6.2f} TF", flush=True)
    pass
    del basis
    torch.cuda.empty_cache()
