"""K3 (ERI sweep), K5 (closed-form per-point features), K6 (XC quadrature) and the Fock glue vs the CPU oracle."""
import math

import pytest
import torch

import oracle
from graddft_b200 import ops
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu

RTOL = 1e-11
F64 = torch.float64


def relerr(a, b):
    return float((a.cpu() - b).abs().max() / (b.abs().max() + 1e-300))


def grid_quantities(N, seed, zero_frac=0.01):
    """rho >= 0 spanning many decades, with a fraction of exactly-zero / sub-clip rows to hit the guards."""
    g = torch.Generator().manual_seed(seed)
    rho = torch.exp(-12.0 * torch.rand(N, 2, generator=g, dtype=F64)) * 3.0
    grho = torch.randn(N, 2, 3, generator=g, dtype=F64) * rho[:, :, None] ** (4.0 / 3.0)
    tau = torch.rand(N, 2, generator=g, dtype=F64) * rho ** (5.0 / 3.0) * 3.0
    lapl = torch.randn(N, 2, generator=g, dtype=F64) * rho
    k = max(1, int(N * zero_frac))
    rho[:k] = 0.0
    grho[:k] = 0.0
    tau[:k] = 0.0
    lapl[:k] = 0.0
    rho[k:2 * k, 0] = 1e-31          # below the clip in one channel only
    rho[2 * k:3 * k, 1] = 0.0        # fully polarised rows
    rho[3 * k:4 * k] = 1e-33
    return rho, grho, tau, lapl


ORACLE_PW = {
    "LSDA_X": lambda r, g, t, l: oracle.lsda_x_e(r).unsqueeze(1),
    "B88_X": lambda r, g, t, l: oracle.b88_x_e(r, g).unsqueeze(1),
    "VWN_C": lambda r, g, t, l: oracle.vwn_c_e(r).unsqueeze(1),
    "LYP_C": lambda r, g, t, l: oracle.lyp_c_e(r, g, l).unsqueeze(1),
    "PW92_C": lambda r, g, t, l: oracle.pw92_c_e(r).unsqueeze(1),
    "B3LYP_SET": lambda r, g, t, l: oracle.b3lyp_exhf_densities(r, g, l),
    "B88_SET": lambda r, g, t, l: torch.stack((oracle.lsda_x_e(r), oracle.b88_x_e(r, g)), dim=1),
    "DM21_INPUTS": lambda r, g, t, l: oracle.dm21_coefficient_inputs(r, g, t),
    "DM21_LDA": lambda r, g, t, l: oracle.dm21_densities(r, g, t, "LDA"),
    "DM21_GGA": lambda r, g, t, l: oracle.dm21_densities(r, g, t, "GGA"),
    "DM21_MGGA": lambda r, g, t, l: oracle.dm21_densities(r, g, t, "MGGA"),
    "FEAT_LDA": lambda r, g, t, l: oracle.mgga_feature_densities(r, g, t, "LDA"),
    "FEAT_GGA": lambda r, g, t, l: oracle.mgga_feature_densities(r, g, t, "GGA"),
    "FEAT_MGGA": lambda r, g, t, l: oracle.mgga_feature_densities(r, g, t, "MGGA"),
}
NEEDS = {  # (grad, tau, lapl)
    "LSDA_X": (0, 0, 0), "B88_X": (1, 0, 0), "VWN_C": (0, 0, 0), "LYP_C": (1, 0, 1), "PW92_C": (0, 0, 0),
    "B3LYP_SET": (1, 0, 1), "B88_SET": (1, 0, 0), "DM21_INPUTS": (1, 1, 0), "DM21_LDA": (0, 0, 0), "DM21_GGA": (1, 0, 0),
    "DM21_MGGA": (1, 1, 0), "FEAT_LDA": (0, 0, 0), "FEAT_GGA": (1, 0, 0), "FEAT_MGGA": (1, 1, 0),
}


@pytest.mark.parametrize("name", list(ORACLE_PW))
@pytest.mark.parametrize("zero_frac", [0.0, 0.01])
def test_pointwise_forward_and_vjp(cuda_device, name, zero_frac):
    N = 4099
    rho, grho, tau, lapl = grid_quantities(N, 1984, zero_frac)
    ng, nt, nl = NEEDS[name]
    dev = cuda_device
    # ---- oracle value + VJP through torch-CPU autograd
    leaves = [t.clone().requires_grad_(True) for t in (rho, grho, tau, lapl)]
    ref = ORACLE_PW[name](*leaves)
    cot = torch.randn(ref.shape, generator=torch.Generator().manual_seed(7), dtype=F64)
    used = [leaves[0]] + [x for x, f in zip(leaves[1:], (ng, nt, nl)) if f]
    ref_grads = torch.autograd.grad((ref * cot).sum(), used, allow_unused=True)
    # ---- kernel
    dl = [rho.to(dev).requires_grad_(True), grho.to(dev).requires_grad_(True) if ng else None,
          tau.to(dev).requires_grad_(True) if nt else None, lapl.to(dev).requires_grad_(True) if nl else None]
    out = ops.pointwise(name, dl[0], dl[1], dl[2], dl[3])
    assert out.shape == ref.shape
    finite = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(out.cpu()), finite), "NaN/inf pattern differs from the reference formulas"
    scale = ref[finite].abs().max()
    assert float((out.cpu()[finite] - ref.detach()[finite]).abs().max() / scale) < RTOL
    used_d = [x for x in dl if x is not None]
    got = torch.autograd.grad((out * cot.to(dev)).sum(), used_d)
    for gg, rg in zip(got, ref_grads):
        rg = torch.zeros_like(gg.cpu()) if rg is None else rg
        fin = torch.isfinite(rg)
        # reverse-mode autodiff of the reference formulas gives NaN (0 * inf through an unselected where
        # branch) only at exactly-zero densities; the kernel returns the finite forward-mode derivative there
        assert bool(torch.isfinite(gg).all()), "kernel VJP must be NaN-free"
        if fin.any():
            sc = rg[fin].abs().max() + 1e-300
            err = (gg.cpu()[fin] - rg[fin]).abs()
            assert bool((err <= 1e-9 * rg[fin].abs() + 1e-12 * sc).all())


@pytest.mark.parametrize("name", list(ORACLE_PW))
def test_pointwise_second_order(cuda_device, name):
    """VJP of the VJP (gdft_pointwise_bwd2) against torch-CPU double backward through the oracle formulas: the
    cotangents of (inputs, out_bar) for random cotangents U of the first-order result.  This is what differentiating
    V_xc once more -- training through the SCF loop, grad_dft/evaluate.py:917-1038 under jax.grad -- needs."""
    N = 2053
    rho, grho, tau, lapl = grid_quantities(N, 1993, 0.0)
    for t in (rho, grho, tau, lapl):  # regular rows only: second derivatives do not exist at exactly-zero / fully polarised points
        t[:4] = t[4:8]
    ng, nt, nl = NEEDS[name]
    dev = cuda_device

    def run(device, pw):
        leaves = [rho.to(device).requires_grad_(True), grho.to(device).requires_grad_(True) if ng else None,
                  tau.to(device).requires_grad_(True) if nt else None, lapl.to(device).requires_grad_(True) if nl else None]
        out = pw(*leaves)
        g2 = torch.Generator().manual_seed(11)
        cot = torch.randn(out.shape, generator=g2, dtype=F64).to(device).requires_grad_(True)
        used = [x for x in leaves if x is not None]
        first = torch.autograd.grad((out * cot).sum(), used, create_graph=True)
        s = 0.0
        for f in first:
            s = s + (f * torch.randn(f.shape, generator=g2, dtype=F64).to(device)).sum()
        second = torch.autograd.grad(s, used + [cot], allow_unused=True)
        return [None if t is None else t.detach().cpu() for t in second], first

    ref, _ = run("cpu", lambda r, g, t, l: ORACLE_PW[name](r, g if g is not None else grho, t if t is not None else tau, l if l is not None else lapl))
    got, first = run(dev, lambda r, g, t, l: ops.pointwise(name, r, g, t, l))
    for a, b in zip(got, ref):
        b = torch.zeros_like(a) if b is None else b
        a = torch.zeros_like(b) if a is None else a
        fin = torch.isfinite(b)
        assert bool(torch.isfinite(a[fin]).all())
        sc = b[fin].abs().max() + 1e-300
        assert bool(((a[fin] - b[fin]).abs() <= 1e-7 * b[fin].abs() + 1e-11 * sc).all()), name
    # third order is not bound and must fail loudly
    with pytest.raises(RuntimeError):
        r3 = rho.to(dev).requires_grad_(True)
        o = ops.pointwise("LSDA_X", r3)
        (g1,) = torch.autograd.grad(o.sum(), r3, create_graph=True)
        (g2_,) = torch.autograd.grad(g1.sum(), r3, create_graph=True)
        torch.autograd.grad(g2_.sum(), r3)


@pytest.mark.parametrize("n", [5, 12, 43, 64, 97])
def test_eri_sweep(cuda_device, n):
    mol = synthetic_molecule(64, n, seed=1993, with_eri=True)
    eri, D = mol["rep_tensor"], mol["rdm1"]
    P = D.sum(dim=0)
    dev = cuda_device
    eri_d, P_d = eri.to(dev), P.to(dev)
    J = ops.coulomb_j(P_d, eri_d)
    assert relerr(J, oracle.coulomb_potential(P, eri)) < RTOL
    J2, EJ = ops.coulomb_j_and_energy(P_d, eri_d)
    assert torch.equal(J2, J)
    assert abs(float(EJ) - float(oracle.coulomb_energy(P, eri))) < 1e-11 * abs(float(oracle.coulomb_energy(P, eri)))
    K = ops.coulomb_k(P_d, eri_d)
    assert relerr(K, torch.einsum("pqrt,qt->pr", eri, P)) < RTOL
    # J and K from ONE pass over the tensor (gdft_eri_jk with K requested); run-to-run reproducible
    Jf, Kf = ops.coulomb_jk(P_d, eri_d)
    assert relerr(Jf, oracle.coulomb_potential(P, eri)) < RTOL and torch.equal(Kf, K)
    assert all(torch.equal(a, b) for a, b in zip(ops.coulomb_jk(P_d, eri_d), (Jf, Kf)))
    # transpose sweep on a NON-symmetric tensor pins the index pairing
    g = torch.Generator().manual_seed(3)
    eri_ns = torch.randn(n, n, n, n, generator=g, dtype=F64)
    Jbar = torch.randn(n, n, generator=g, dtype=F64)
    Pq = P.clone().requires_grad_(True)
    (ref,) = torch.autograd.grad((oracle.coulomb_potential(Pq, eri_ns) * Jbar).sum(), Pq)
    Pd = P_d.clone().requires_grad_(True)
    (got,) = torch.autograd.grad((ops.coulomb_j(Pd, eri_ns.to(dev)) * Jbar.to(dev)).sum(), Pd)
    assert relerr(got, ref) < RTOL
    assert relerr(ops.coulomb_j(P_d, eri_ns.to(dev)), oracle.coulomb_potential(P, eri_ns)) < RTOL
    # the same for the J+K sweep: values, the VJP through both outputs and the VJP of the VJP, on the non-symmetric tensor
    Kbar = torch.randn(n, n, generator=g, dtype=F64)
    Pq = P.clone().requires_grad_(True)
    Jr, Kr = oracle.coulomb_potential(Pq, eri_ns), torch.einsum("pqrt,qt->pr", eri_ns, Pq)
    (ref,) = torch.autograd.grad((Jr * Jbar).sum() + (Kr * Kbar).sum(), Pq)
    (ref_k,) = torch.autograd.grad((torch.einsum("pqrt,qt->pr", eri_ns, Pq) * Kbar).sum(), Pq)
    Pd = P_d.clone().requires_grad_(True)
    Jg, Kg = ops.coulomb_jk(Pd, eri_ns.to(dev))
    assert relerr(Jg.detach(), Jr.detach()) < RTOL and relerr(Kg.detach(), Kr.detach()) < RTOL
    (got,) = torch.autograd.grad((Jg * Jbar.to(dev)).sum() + (Kg * Kbar.to(dev)).sum(), Pd)
    assert relerr(got, ref) < RTOL
    Pd = P_d.clone().requires_grad_(True)
    (got_k,) = torch.autograd.grad((ops.coulomb_k(Pd, eri_ns.to(dev)) * Kbar.to(dev)).sum(), Pd)
    assert relerr(got_k, ref_k) < RTOL
    # K is linear in P: the VJP of its VJP is the sweep itself
    Kb = Kbar.to(dev).clone().requires_grad_(True)
    Pd = P_d.clone().requires_grad_(True)
    (v,) = torch.autograd.grad((ops.coulomb_k(Pd, eri_ns.to(dev)) * Kb).sum(), Pd, create_graph=True)
    U = torch.randn(n, n, generator=g, dtype=F64)
    (vv,) = torch.autograd.grad((v * U.to(dev)).sum(), Kb)
    assert relerr(vv, torch.einsum("pqrt,qt->pr", eri_ns, U)) < RTOL


@pytest.mark.parametrize("n,world", [(12, 2), (43, 3), (64, 8)])
def test_eri_row_sharded_sweep(cuda_device, n, world):
    """Row-sharded J: gathered row blocks are bitwise the unsharded sweep, and the blocks'
    transposed sweeps sum to the unsharded VJP."""
    from graddft_b200 import distributed as gdist

    g = torch.Generator().manual_seed(5)
    eri = torch.randn(n, n, n, n, generator=g, dtype=F64)
    P = torch.randn(n, n, generator=g, dtype=F64)
    Jbar = torch.randn(n, n, generator=g, dtype=F64)
    dev = cuda_device
    J_full = ops.coulomb_j(P.to(dev), eri.to(dev))
    Pl = P.to(dev).clone().requires_grad_(True)
    (vjp_full,) = torch.autograd.grad((ops.coulomb_j(Pl, eri.to(dev)) * Jbar.to(dev)).sum(), Pl)
    rows, vjp = [], torch.zeros(n, n, dtype=F64, device=dev)
    for r in range(world):
        part = gdist.shard_molecule_tensors({"weights": torch.zeros(8, dtype=F64), "rep_tensor": eri}, r, world, shard_eri=True)
        blk = part["rep_tensor"].to(dev)
        Pl = P.to(dev).clone().requires_grad_(True)
        Jr = ops.coulomb_j_rows(Pl, blk)
        rows.append(Jr.detach())
        r0 = part["eri_row0"]
        if blk.shape[0]:
            (v,) = torch.autograd.grad((Jr * Jbar.to(dev).reshape(-1)[r0:r0 + blk.shape[0]]).sum(), Pl)
            vjp += v
    assert torch.equal(torch.cat(rows).reshape(n, n), J_full)
    assert relerr(vjp, vjp_full.cpu()) < RTOL
    assert relerr(J_full, oracle.coulomb_potential(P, eri)) < RTOL


@pytest.mark.parametrize("N,F,crows", [(1, 1, 1), (1000, 5, 1), (4097, 3, 4097), (300001, 5, 1), (70000, 20, 70000)])
def test_xc_integrate(cuda_device, N, F, crows):
    g = torch.Generator().manual_seed(N + F)
    c = torch.randn(crows, F, generator=g, dtype=F64)
    d = torch.randn(N, F, generator=g, dtype=F64) * torch.exp(-20 * torch.rand(N, 1, generator=g, dtype=F64))
    w = torch.rand(N, generator=g, dtype=F64)
    k = max(1, N // 50)
    d[:k] = 1e-32       # |e| below the clip
    w[k:2 * k] = 1e-31  # weight below the clip
    dev = cuda_device
    cl, dl_ = c.clone().requires_grad_(True), d.clone().requires_grad_(True)
    ref = oracle.xc_energy(cl, dl_, w)
    rc, rd = torch.autograd.grad(ref, (cl, dl_))
    cd, dd = c.to(dev).requires_grad_(True), d.to(dev).requires_grad_(True)
    E = ops.xc_integrate(cd, dd, w.to(dev))
    assert abs(float(E) - float(ref)) <= 1e-12 * max(1.0, abs(float(ref)))
    gc, gd = torch.autograd.grad(E, (cd, dd))
    assert relerr(gc, rc) < RTOL and relerr(gd, rd) < RTOL
    # run-to-run bitwise reproducibility
    assert float(ops.xc_integrate(cd, dd, w.to(dev))) == float(E)


def test_fock_glue(cuda_device):
    n = 37
    g = torch.Generator().manual_seed(11)
    h = torch.randn(n, n, generator=g, dtype=F64)
    J = torch.randn(n, n, generator=g, dtype=F64)
    Db = torch.randn(2, n, n, generator=g, dtype=F64)
    Db[0, 0, 1] = 1e-31 - h[0, 1] - J[0, 1]  # lands under the clip before symmetrisation
    dev = cuda_device
    x = oracle.abs_clip(h + J + Db)
    ref = oracle.abs_clip(0.5 * (x + x.transpose(1, 2)))
    got = ops.fock_assemble(h.to(dev), J.to(dev), Db.to(dev))
    assert torch.equal(got.cpu(), ref)
    V = torch.randn(2, n, n, generator=g, dtype=F64)
    ref2 = oracle.abs_clip(ref + (V + V.transpose(1, 2)))  # train.py:205: fock += V + V^T
    got2 = ops.fock_add_sym_(got.clone(), V.to(dev))
    assert torch.equal(got2.cpu(), ref2)


@pytest.mark.parametrize("N,W,with_res,with_ybias", [(1000, 256, True, True), (333, 16, True, False), (77, 512, False, True),
                                                      (4097, 64, True, True), (5000, 256, True, False)])
def test_residual_layernorm_elu(cuda_device, N, W, with_res, with_ybias):
    """Row f2: elu(LayerNorm(y + ybias + res) * scale + bias) (grad_dft/functional.py:809-819, the Dense bias folded in) fused,
    value and first-order cotangents of (y, ybias, res, scale, bias) against the same composite in torch-CPU float64."""
    g = torch.Generator().manual_seed(21)
    y = torch.randn(N, W, generator=g, dtype=F64) * 2.0
    ybias = 0.5 * torch.randn(W, generator=g, dtype=F64)
    res = torch.randn(N, W, generator=g, dtype=F64)
    scale = 1.0 + 0.3 * torch.randn(W, generator=g, dtype=F64)
    bias = 0.2 * torch.randn(W, generator=g, dtype=F64)
    cot = torch.randn(N, W, generator=g, dtype=F64)

    def composite(y, ybias, res, scale, bias):
        z = y + (ybias if with_ybias else 0.0) + (res if with_res else 0.0)
        mu = z.mean(dim=-1, keepdim=True)
        var = ((z - mu) ** 2).mean(dim=-1, keepdim=True)
        return torch.nn.functional.elu((z - mu) * torch.rsqrt(var + 1e-6) * scale + bias)

    use = [True, with_ybias, with_res, True, True]
    leaves = [t.clone().requires_grad_(True) for t in (y, ybias, res, scale, bias)]
    ref = composite(*leaves)
    ref_g = torch.autograd.grad((ref * cot).sum(), [l for l, u in zip(leaves, use) if u])
    dl = [t.to(cuda_device).requires_grad_(True) for t in (y, ybias, res, scale, bias)]
    with ops.first_order_build():
        assert ops.residual_layernorm_elu_supported(dl[0])
        out = ops.residual_layernorm_elu(dl[0], dl[2] if with_res else None, dl[3], dl[4], 1e-6, ybias=dl[1] if with_ybias else None)
    assert relerr(out, ref.detach()) < 1e-13
    got = torch.autograd.grad((out * cot.to(cuda_device)).sum(), [l for l, u in zip(dl, use) if u])
    for a, b in zip(got, ref_g):
        assert relerr(a, b) < 1e-12
    assert not ops.residual_layernorm_elu_supported(dl[0])  # outside a first-order build the composite is used


@pytest.mark.parametrize("n", [1, 2, 3, 5, 12, 31, 43, 44, 45, 63, 64, 65, 77, 90])
def test_sym_eigh_jacobi(cuda_device, n):
    """Row f1: the small symmetric eigenproblem of the SCF iteration (utils/eigenproblem.py:26-106) against LAPACK on the
    CPU: ascending eigenvalues, orthonormal eigenvectors, small residual; degenerate and diagonal inputs included."""
    g = torch.Generator().manual_seed(100 + n)
    A = torch.randn(4, n, n, generator=g, dtype=F64)
    A = 0.5 * (A + A.transpose(1, 2))
    A[1] = torch.diag(torch.randn(n, generator=g, dtype=F64))                      # already diagonal
    u = torch.randn(n, 1, generator=g, dtype=F64)
    A[2] = torch.eye(n, dtype=F64) * 3.0 + (u @ u.T if n > 1 else 0.0)             # (n-1)-fold degenerate
    A[3] = A[3] * 1e-8 + torch.diag(torch.linspace(-50.0, 50.0, n, dtype=F64))     # widely spread, nearly diagonal
    assert ops.sym_eigh_supported(A.to(cuda_device))
    w, V = ops.sym_eigh(A.to(cuda_device))
    w, V = w.cpu(), V.cpu()
    w_ref = torch.linalg.eigvalsh(A)
    scale = A.abs().amax(dim=(1, 2)).clamp_min(1e-300)
    assert bool(((w - w_ref).abs().amax(dim=1) <= 1e-13 * scale * max(n, 4)).all())
    assert bool((w[:, 1:] >= w[:, :-1]).all())
    eye = torch.eye(n, dtype=F64)
    assert float((V.transpose(1, 2) @ V - eye).abs().max()) < 1e-13 * max(n, 4)
    resid = (A @ V - V * w.unsqueeze(1)).abs().amax(dim=(1, 2))
    assert bool((resid <= 1e-13 * scale * max(n, 4)).all())
    assert not ops.sym_eigh_supported(torch.zeros(2, 400, 400, dtype=F64, device=cuda_device))


@pytest.mark.parametrize("m,n", [(10, 43), (10, 7), (3, 64), (10, 264)])
def test_diis_reductions(cuda_device, m, n):
    """The two CDIIS reductions (grad_dft/evaluate.py:1165, 1198) against the einsums they replace."""
    g = torch.Generator().manual_seed(m * 1000 + n)
    e = torch.randn(m, 2, n, n, generator=g, dtype=F64)
    e[m - 1] = 0.0  # a dead ring slot, as in the first cycles
    f = torch.randn(m, 2, n, n, generator=g, dtype=F64)
    x = torch.randn(2, m, generator=g, dtype=F64)
    G = ops.diis_gram(e.to(cuda_device))
    assert relerr(G, torch.einsum("iskl,jskl->sij", e, e)) < 1e-13
    assert torch.equal(G, G.transpose(1, 2))
    assert relerr(ops.diis_combine(x.to(cuda_device), f.to(cuda_device)), torch.einsum("si,isjk->sjk", x, f)) < 1e-13


def test_abs_clip_kernel(cuda_device):
    """abs_clip (grad_dft/molecule.py:687-689): value, VJP and the VJP of the VJP against the torch composite."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1000, 3, generator=g, dtype=F64)
    x[::7] *= 1e-31
    x[5, 1] = 0.0
    x[6, 2] = float("nan")
    cot = torch.randn(1000, 3, generator=g, dtype=F64)
    ref_in = x.clone().requires_grad_(True)
    ref = torch.where(ref_in.abs() > 1e-30, ref_in, torch.zeros_like(ref_in))
    (ref_g,) = torch.autograd.grad((ref * cot).sum(), ref_in)
    xin = x.to(cuda_device).requires_grad_(True)
    from graddft_b200.molecule import abs_clip
    out = abs_clip(xin, 1e-30)
    assert torch.equal(out.detach().cpu(), ref.detach())
    c = cot.to(cuda_device).requires_grad_(True)
    (gx,) = torch.autograd.grad((out * c).sum(), xin, create_graph=True)
    assert torch.equal(gx.detach().cpu(), ref_g)
    (gc,) = torch.autograd.grad(gx.sum(), c)   # second order: d/dc of mask * c
    assert torch.equal(gc.cpu(), (x.abs() > 1e-30).to(F64))


@pytest.mark.parametrize("cycle", [0, 3, 9, 10, 14])
def test_diis_bordered_matrix(cuda_device, cycle):
    """gdft_diis_matrix against the assembly of grad_dft/evaluate.py:1167-1181 as the host mirror spells it out."""
    from graddft_b200.evaluate import JittableDiis
    m, n = 10, 13
    g = torch.Generator().manual_seed(cycle)
    e = torch.randn(m, 2, n, n, generator=g, dtype=F64)
    e[min(cycle, m - 1) + 1:] = 0.0
    B = ops.diis_matrix(e.to(cuda_device), cycle).cpu()
    G = torch.einsum("iskl,jskl->sij", e, e)
    ref = torch.zeros((2, m + 1, m + 1), dtype=F64)
    ref[:, 1:, 1:] = G
    live = (torch.arange(m) <= cycle).to(F64)
    ref[:, 0, 1:] = live
    ref[:, 1:, 0] = live
    idx = torch.arange(1, m + 1)
    ref[:, idx, idx] = torch.where(live.bool(), torch.diagonal(G, dim1=1, dim2=2), torch.ones_like(live))
    assert relerr(B, ref) < 1e-13
    d = JittableDiis(torch.eye(n, dtype=F64, device=cuda_device), torch.eye(n, dtype=F64, device=cuda_device), m)
    x_fused = d.cdiis_minimize(e.to(cuda_device), cycle)
    x_ref = (torch.linalg.inv(ref) @ torch.tensor([1.0] + [0.0] * m, dtype=F64))[:, 1:]
    assert relerr(x_fused, x_ref) < 1e-9


@pytest.mark.parametrize("n", [2, 5, 12, 43, 44, 63, 64])
def test_sym_eigh_warm_start(cuda_device, n):
    """gdft_sym_eigh_warm: with the eigenvectors of a nearby matrix as V0 the decomposition of A is the same (eigenvalues to
    1e-13, A V = V diag(w), V orthogonal), also when V0 already diagonalises A exactly and when it is unrelated to A."""
    g = torch.Generator().manual_seed(100 + n)
    A = torch.randn(2, n, n, generator=g, dtype=F64)
    A = A + A.transpose(1, 2)
    P = torch.randn(2, n, n, generator=g, dtype=F64)
    A2 = A + 1e-3 * (P + P.transpose(1, 2))
    w_ref = torch.linalg.eigvalsh(A2)
    Ad, A2d = A.to(cuda_device), A2.to(cuda_device)
    _, V_prev = ops.sym_eigh(Ad)
    Q, _ = torch.linalg.qr(torch.randn(2, n, n, generator=g, dtype=F64))
    for V0 in (V_prev, ops.sym_eigh(A2d)[1], Q.to(cuda_device)):
        w, V = ops.sym_eigh(A2d, V0)
        assert float((w.cpu() - w_ref).abs().max()) < 1e-12 * float(w_ref.abs().max())
        resid = (A2d @ V - V * w.unsqueeze(-2)).abs().max()
        assert float(resid) < 1e-12 * float(w_ref.abs().max())
        eye = torch.eye(n, dtype=F64, device=cuda_device)
        assert float((V.transpose(1, 2) @ V - eye).abs().max()) < 1e-12


@pytest.mark.parametrize("n", [96, 264])
def test_refine_eigh(cuda_device, n):
    """evaluate.refine_eigh (Ogita-Aishima refinement of the previous cycle's eigenvectors, n beyond the Jacobi kernel):
    accepted results are eigen-decompositions to working accuracy -- also with an exactly degenerate pair --, and starts that
    are too far (an unrelated basis; a perturbation larger than the gaps) are refused, never answered wrongly."""
    from graddft_b200.evaluate import refine_eigh
    g = torch.Generator().manual_seed(n)
    Q, _ = torch.linalg.qr(torch.randn(2, n, n, generator=g, dtype=F64))
    lam = torch.linspace(-4.0, 5.0, n, dtype=F64).repeat(2, 1) + 0.01 * torch.rand(2, n, generator=g, dtype=F64)
    lam[:, 5] = lam[:, 4]  # an exactly double eigenvalue
    A = (Q * lam.unsqueeze(-2)) @ Q.transpose(1, 2)
    A = 0.5 * (A + A.transpose(1, 2))
    P = torch.randn(2, n, n, generator=g, dtype=F64)
    P = P + P.transpose(1, 2)
    lam2 = lam.clone()
    lam2[:, 5] += 0.03  # separated spectrum for the perturbed case
    B = (Q * lam2.unsqueeze(-2)) @ Q.transpose(1, 2)
    B = 0.5 * (B + B.transpose(1, 2)) + 1e-6 * P
    eye = torch.eye(n, dtype=F64, device=cuda_device)
    Qd = Q.to(cuda_device)
    for M in (A, B):
        ref = torch.linalg.eigvalsh(M)
        Md = M.to(cuda_device)
        w, V = refine_eigh(Md, Qd)
        assert w is not None
        scale = float(ref.abs().max())
        assert float((w.cpu() - ref).abs().max()) < 1e-12 * scale
        assert float((Md @ V - V * w.unsqueeze(-2)).abs().max()) < 1e-11 * scale
        assert float((V.transpose(1, 2) @ V - eye).abs().max()) < 1e-12
    far, _ = torch.linalg.qr(torch.randn(2, n, n, generator=g, dtype=F64))
    for M, X0 in ((B, far), (B + 1e-2 * P, Q)):
        w, V = refine_eigh(M.to(cuda_device), X0.to(cuda_device))
        if w is not None:  # a refusal is the expected outcome; an answer must still be a decomposition
            Md = M.to(cuda_device)
            assert float((Md @ V - V * w.unsqueeze(-2)).abs().max()) < 1e-11 * float(w.abs().max())
    assert refine_eigh(B.to(cuda_device), far.to(cuda_device))[0] is None


@pytest.mark.parametrize("n", [65, 66, 90, 96, 128, 161, 200, 264, 265, 320])
def test_sym_eigh_cluster(cuda_device, n):
    """Row f1 beyond n = 64 (csrc/eigh_cluster.cu: one 8-CTA cluster per matrix, one-sided Jacobi on (A + sigma I) V0) against
    LAPACK on the CPU at 1e-12: random, diagonal, (n-1)-fold degenerate and nearly diagonal inputs; the status word; a
    warm start from the eigenvectors of a nearby matrix, from the exact eigenvectors and from an unrelated orthogonal basis."""
    g = torch.Generator().manual_seed(100 + n)
    A = torch.randn(4, n, n, generator=g, dtype=F64)
    A = 0.5 * (A + A.transpose(1, 2))
    A[1] = torch.diag(torch.randn(n, generator=g, dtype=F64))
    u = torch.randn(n, 1, generator=g, dtype=F64)
    A[2] = torch.eye(n, dtype=F64) * 3.0 + u @ u.T
    A[3] = A[3] * 1e-8 + torch.diag(torch.linspace(-50.0, 50.0, n, dtype=F64))
    eye = torch.eye(n, dtype=F64)

    def check(Ah, w, V, tol=1e-12):
        w, V = w.cpu(), V.cpu()
        w_ref = torch.linalg.eigvalsh(Ah)
        rho = w_ref.abs().amax(dim=-1).clamp_min(1e-300)
        assert bool(((w - w_ref).abs().amax(dim=-1) <= tol * rho).all()), float(((w - w_ref).abs().amax(dim=-1) / rho).max())
        assert bool((w[..., 1:] >= w[..., :-1]).all())
        assert float((V.transpose(-1, -2) @ V - eye).abs().max()) < tol
        resid = (Ah @ V - V * w.unsqueeze(-2)).abs().amax(dim=(-1, -2))
        assert bool((resid <= tol * rho).all()), float((resid / rho).max())

    Ad = A.to(cuda_device)
    assert ops.sym_eigh_supported(Ad)
    info = torch.full((4,), -99, dtype=torch.int32, device=cuda_device)
    w, V = ops.sym_eigh(Ad, info=info)
    check(A, w, V)
    cold = info.cpu()
    assert bool((cold >= 1).all()) and bool((cold <= 20).all()), cold
    assert int(cold[1]) == 1  # a diagonal matrix: one sweep without a single rotation
    # warm starts
    P = torch.randn(2, n, n, generator=g, dtype=F64)
    A2 = A[:2] + 1e-4 * (P + P.transpose(1, 2))
    A2[1] = A[0] + 1e-7 * (P[1] + P[1].T)
    A2d = A2.to(cuda_device)
    V_prev = torch.stack([V[0], V[0]])
    info2 = torch.full((2,), -99, dtype=torch.int32, device=cuda_device)
    w2, V2 = ops.sym_eigh(A2d, V_prev, info=info2)
    check(A2, w2, V2)
    assert int(info2[0]) < int(cold[0]) and int(info2[1]) <= int(info2[0]), (info2, cold)
    w3, V3 = ops.sym_eigh(A2d, V2, info=info2)  # the exact eigenvectors: nothing to rotate
    check(A2, w3, V3)
    assert int(info2.max()) <= 2
    Q, _ = torch.linalg.qr(torch.randn(2, n, n, generator=g, dtype=F64))
    w4, V4 = ops.sym_eigh(A2d, Q.to(cuda_device), info=info2)
    check(A2, w4, V4)
    # a non-finite input is reported, not hidden
    bad = Ad[:1].clone()
    bad[0, 3, 5] = float("nan")
    ops.sym_eigh(bad, info=info2)
    assert int(info2[0]) == -2


def test_sym_eigh_status_word_small(cuda_device):
    """The one-CTA kernels report their sweep count too (ADVICE: no silent non-convergence inside a captured graph)."""
    g = torch.Generator().manual_seed(5)
    A = torch.randn(3, 43, 43, generator=g, dtype=F64)
    A = (A + A.transpose(1, 2)).to(cuda_device)
    info = torch.full((3,), -99, dtype=torch.int32, device=cuda_device)
    ops.sym_eigh(A, info=info)
    assert bool((info >= 1).all()) and bool((info <= 12).all()), info
    A[1, 0, 0] = float("inf")
    ops.sym_eigh(A, info=info)
    assert int(info[1]) == -2 and int(info[0]) >= 1


@pytest.mark.parametrize("n", [13, 43, 64])
def test_sym_eigh_warm_start_reorthogonalises_its_start(cuda_device, n):
    """The one-CTA Jacobi kernel returns V0 V'; chained over many SCF cycles the product would drift from orthogonality.
    The kernel applies one Newton-Schulz step to V0 first: a start that is orthogonal only to 1e-8 still yields an
    orthogonal decomposition, and a 200-cycle chain stays at working accuracy (no periodic cold start)."""
    g = torch.Generator().manual_seed(n)
    A = torch.randn(2, n, n, generator=g, dtype=torch.float64)
    A = (A + A.transpose(1, 2)).to(cuda_device)
    eye = torch.eye(n, dtype=torch.float64, device=cuda_device)
    _, V = ops.sym_eigh(A)
    V_bad = V + 1e-8 * torch.randn(V.shape, generator=g, dtype=torch.float64).to(cuda_device)  # defect 1e-7: one step leaves its square
    w, V2 = ops.sym_eigh(A, V_bad)
    assert float((V2.transpose(1, 2) @ V2 - eye).abs().max()) < 1e-12
    assert float((V2.transpose(1, 2) @ A @ V2 - torch.diag_embed(w)).abs().max()) < 1e-11 * float(A.abs().max()) * n
    Vk = V
    for k in range(200):
        P = torch.randn(2, n, n, generator=g, dtype=torch.float64).to(cuda_device)
        Ak = A + 1e-3 * (P + P.transpose(1, 2))
        wk, Vk = ops.sym_eigh(Ak, Vk)
    assert float((Vk.transpose(1, 2) @ Vk - eye).abs().max()) < 1e-13 * n
    w_ref = torch.linalg.eigvalsh(Ak)
    assert float((wk - w_ref).abs().max()) < 1e-11 * float(A.abs().max()) * n
