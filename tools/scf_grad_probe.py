"""Development probe: autograd vs finite differences of E_scf w.r.t. neural-functional params."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import oracle
import graddft_b200 as gd
from graddft_b200.synthetic import synthetic_molecule
F64 = torch.float64
dev = torch.device("cuda:0")

def gapped(N, n, seed):
    mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
    g = torch.Generator().manual_seed(seed)
    mol["h1e"] = torch.diag(torch.linspace(-8.0, 8.0, n, dtype=F64)) + 0.05 * mol["h1e"]
    mol["rep_tensor"] = 0.05 * mol["rep_tensor"]
    mol["s1e"] = torch.eye(n, dtype=F64) + 0.2 * (mol["s1e"] - torch.eye(n, dtype=F64))
    return mol

mol = gapped(1200, 8, 1984)
m = gd.molecule_from_tensors(mol, dev)
import sys as _s
if "--hybrid" in _s.argv:
    fun = gd.DM21(layer_widths=(8, 8))
    flat = oracle.dm21_mlp_init(width=8, n_layers=2, seed=3)
else:
    fun = gd.DM21(layer_widths=(8, 8), nograd_densities=None, densitygrads=None, combine_densities=None, nograd_coefficient_inputs=None,
                  coefficient_input_grads=None, combine_inputs=None, local_features=1, needs_omegas=None)
    flat = fun.generate_DM21_weights(n_input_features=7, seed=3)
gen = torch.Generator().manual_seed(5)
direction = {k: torch.randn(v.shape, generator=gen, dtype=F64).to(dev) for k, v in flat.items()}
for name, mk in (("simple", lambda c: gd.diff_simple_scf_loop(fun, cycles=c)), ("diis", lambda c: gd.diff_scf_loop(fun, cycles=c))):
    for cycles in (0, 1, 2, 4):
        loop = mk(cycles)
        params = {k: v.to(dev).requires_grad_(True) for k, v in flat.items()}
        e = loop(params, m).energy
        grads = torch.autograd.grad(e, list(params.values()), allow_unused=True)
        slope = sum(float((g * direction[k]).sum()) for g, k in zip(grads, params) if g is not None)
        fds = []
        for h in (1e-4, 1e-5, 1e-6):
            with torch.no_grad():
                ep = loop({k: v.to(dev) + h * direction[k] for k, v in flat.items()}, m).energy
                em = loop({k: v.to(dev) - h * direction[k] for k, v in flat.items()}, m).energy
            fds.append(float(ep - em) / (2 * h))
        print(f"{name} cycles={cycles} E={float(e):.10f} autograd={slope:.10e} fd={fds}")
