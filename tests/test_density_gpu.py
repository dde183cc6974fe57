"""K1/K2/K4 parity: density family forward, its transpose and the HF contractions vs the CPU oracle."""
import pytest
import torch

import oracle
from graddft_b200 import ops
from graddft_b200._lib import GDFT_GRAD, GDFT_HF, GDFT_LAPL, GDFT_RHO, GDFT_TAU
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu

# relative to the largest magnitude of the reference tensor; BASELINE.json asks for 1e-7 relative
RTOL = 1e-11

SHAPES = [(1000, 12, 1984), (777, 43, 1993), (2049, 80, 1984), (1500, 264, 1993), (1300, 400, 1984), (513, 7, 1993)]


def relerr(a, b):
    return float((a.cpu() - b).abs().max() / (b.abs().max() + 1e-300))


def make(N, n, seed, dev, W=2, symmetric=True):
    mol = synthetic_molecule(N, n, n_omega=W, seed=seed, with_eri=False, symmetric_rdm1=symmetric)
    basis = ops.PackedBasis(mol["ao"].to(dev), mol["grad_ao"].to(dev), mol["grad_n_ao2"].to(dev), mol["chi"].to(dev))
    return mol, basis


@pytest.mark.parametrize("N,n,seed", SHAPES)
@pytest.mark.parametrize("symmetric", [True, False])
def test_density_forward(cuda_device, N, n, seed, symmetric):
    mol, basis = make(N, n, seed, cuda_device, symmetric=symmetric)
    D = mol["rdm1"]
    flags = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL | GDFT_HF
    rho, grho, tau, lapl, ehf = ops.density_forward(basis, D.to(cuda_device), flags)
    assert relerr(rho, oracle.density(D, mol["ao"])) < RTOL
    assert relerr(grho, oracle.grad_density(D, mol["ao"], mol["grad_ao"])) < RTOL
    assert relerr(tau, oracle.kinetic_density(D, mol["grad_ao"])) < RTOL
    assert relerr(lapl, oracle.lapl_density(D, mol["ao"], mol["grad_ao"], mol["grad_n_ao2"])) < RTOL
    assert relerr(ehf, oracle.HF_energy_density(D, mol["ao"], mol["chi"])) < RTOL
    # subsets must agree bit for bit with the full call (same tiles, same order)
    rho2, grho2, _, _, _ = ops.density_forward(basis, D.to(cuda_device), GDFT_RHO | GDFT_GRAD)
    assert torch.equal(rho2, rho) and torch.equal(grho2, grho)
    tau2 = ops.density_forward(basis, D.to(cuda_device), GDFT_TAU)[2]
    assert torch.equal(tau2, tau)


@pytest.mark.parametrize("N,n,seed", SHAPES)
def test_density_transpose(cuda_device, N, n, seed):
    mol, basis = make(N, n, seed, cuda_device)
    g = torch.Generator().manual_seed(seed + 1)
    rb = torch.randn(N, 2, generator=g, dtype=torch.float64)
    gb = torch.randn(N, 2, 3, generator=g, dtype=torch.float64)
    tb = torch.randn(N, 2, generator=g, dtype=torch.float64)
    lb = torch.randn(N, 2, generator=g, dtype=torch.float64)
    lap_ao = mol["grad_n_ao2"].sum(dim=-1)
    dev = cuda_device
    cases = [
        dict(rho_bar=rb), dict(rho_bar=rb, grho_bar=gb), dict(tau_bar=tb), dict(rho_bar=rb, grho_bar=gb, tau_bar=tb),
        dict(rho_bar=rb, grho_bar=gb, lapl_bar=lb), dict(rho_bar=rb, grho_bar=gb, tau_bar=tb, lapl_bar=lb), dict(lapl_bar=lb),
    ]
    for kw in cases:
        ref = oracle.density_vjp_formula(mol["ao"], mol["grad_ao"], lap_ao, **kw)
        out = ops.density_transpose(basis, **{k: v.to(dev) for k, v in kw.items()})
        assert relerr(out, ref) < RTOL, kw.keys()
    # run-to-run bitwise reproducibility of the split-K reduction
    a = ops.density_transpose(basis, rho_bar=rb.to(dev), grho_bar=gb.to(dev))
    b = ops.density_transpose(basis, rho_bar=rb.to(dev), grho_bar=gb.to(dev))
    assert torch.equal(a, b)


@pytest.mark.parametrize("N,n,seed", SHAPES[:4])
def test_density_vjp_matches_oracle_autograd(cuda_device, N, n, seed):
    """jax.grad-style check: VJP through the bound ops == torch-CPU autograd through the oracle einsums."""
    mol, basis = make(N, n, seed, cuda_device, symmetric=False)
    g = torch.Generator().manual_seed(seed + 2)
    cot = [torch.randn(N, 2, generator=g, dtype=torch.float64), torch.randn(N, 2, 3, generator=g, dtype=torch.float64),
           torch.randn(N, 2, generator=g, dtype=torch.float64), torch.randn(N, 2, generator=g, dtype=torch.float64),
           torch.randn(2, 2, N, generator=g, dtype=torch.float64)]
    D = mol["rdm1"].clone().requires_grad_(True)
    outs = [oracle.density(D, mol["ao"]), oracle.grad_density(D, mol["ao"], mol["grad_ao"]), oracle.kinetic_density(D, mol["grad_ao"]),
            oracle.lapl_density(D, mol["ao"], mol["grad_ao"], mol["grad_n_ao2"]), oracle.HF_energy_density(D, mol["ao"], mol["chi"])]
    (ref,) = torch.autograd.grad(sum((o * c).sum() for o, c in zip(outs, cot)), D)
    Dg = mol["rdm1"].to(cuda_device).requires_grad_(True)
    flags = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL | GDFT_HF
    outs_g = ops.density_forward(basis, Dg, flags)
    (got,) = torch.autograd.grad(sum((o * c.to(cuda_device)).sum() for o, c in zip(outs_g, cot)), Dg)
    assert relerr(got, ref) < RTOL


@pytest.mark.parametrize("N,n,seed", SHAPES[:4])
def test_hf_fock(cuda_device, N, n, seed):
    mol, basis = make(N, n, seed, cuda_device)
    g = torch.randn(2, 2, N, generator=torch.Generator().manual_seed(seed + 3), dtype=torch.float64)
    ref = oracle.HF_fock(mol["chi"], g, mol["ao"])
    out = ops.hf_fock(basis, g.to(cuda_device))
    assert relerr(out, ref) < RTOL


def test_second_order_closure(cuda_device):
    """L and L^T are each other's VJP: double-backward through density_forward stays on the same kernels."""
    N, n = 600, 24
    mol, basis = make(N, n, 1984, cuda_device)
    dev = cuda_device
    D = mol["rdm1"].to(dev).requires_grad_(True)
    w = mol["weights"].to(dev)
    rho = ops.density_forward(basis, D, GDFT_RHO)[0]
    E = (w[:, None] * rho ** 2).sum()
    (g1,) = torch.autograd.grad(E, D, create_graph=True)
    v = torch.randn(2, n, n, generator=torch.Generator().manual_seed(5), dtype=torch.float64)
    (hv,) = torch.autograd.grad((g1 * v.to(dev)).sum(), D)
    Dc = mol["rdm1"].clone().requires_grad_(True)
    Ec = (mol["weights"][:, None] * oracle.density(Dc, mol["ao"]) ** 2).sum()
    (g1c,) = torch.autograd.grad(Ec, Dc, create_graph=True)
    (hvc,) = torch.autograd.grad((g1c * v).sum(), Dc)
    assert relerr(g1.detach(), g1c.detach()) < RTOL
    assert relerr(hv, hvc) < RTOL


@pytest.mark.parametrize("N,n,seed", [(1000, 12, 1984), (777, 43, 1993), (1500, 264, 1993), (1300, 400, 1984), (513, 7, 1993)])
def test_density_forward_row_tile_shapes_agree_bitwise(cuda_device, N, n, seed, monkeypatch):
    """K1's 64-row CTA shape (mid-size grids) against its 128-row shape: the per-row summation order does not depend on
    the tile height, so every output is bitwise the same -- and both match the oracle."""
    mol, basis = make(N, n, seed, cuda_device)
    D = mol["rdm1"].to(cuda_device)
    flags = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL | GDFT_HF
    monkeypatch.setenv("GDFT_FWD_ROWS", "128")
    ref = ops._density_fwd_raw(basis, D, flags)
    monkeypatch.setenv("GDFT_FWD_ROWS", "64")
    got = ops._density_fwd_raw(basis, D, flags)
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    assert relerr(got[0], oracle.density(mol["rdm1"], mol["ao"])) < RTOL
    assert relerr(got[2], oracle.kinetic_density(mol["rdm1"], mol["grad_ao"])) < RTOL


@pytest.mark.parametrize("N,n,W", [(1000, 12, 2), (777, 43, 1), (1500, 264, 2), (900, 80, 3)])
def test_hf_fock_summed_over_omega_inside_the_gemm(cuda_device, N, n, W):
    """gdft_hf_fock_sum: sum_w -1/2 ao^T diag(g[w,s]) chi[w,s] with the omega sum formed in registers (two omegas per GEMM) against
    the per-omega kernel summed afterwards and against the oracle (grad_dft/molecule.py:606-613 + functional.py:714-717)."""
    mol = synthetic_molecule(N, n, n_omega=W, seed=1984 + W, with_eri=False)
    basis = ops.PackedBasis(mol["ao"].to(cuda_device), mol["grad_ao"].to(cuda_device), None, mol["chi"].to(cuda_device))
    g = torch.randn(W, 2, N, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    got = ops.hf_fock_sum(basis, g.to(cuda_device))
    ref = oracle.HF_fock(mol["chi"], g, mol["ao"]).sum(dim=0)
    assert relerr(got, ref) < RTOL
    assert relerr(got, ops.hf_fock(basis, g.to(cuda_device)).sum(dim=0).cpu()) < RTOL
