"""The JAX wrappers of graddft_b200/jax_ffi.py executed WITHOUT JAX (tests/jax_min_shim.py stands in for custom_vjp / ffi_call
and routes every call through the real XLA adapters): values, first-order VJP rules and the second-order closure against the
torch bindings of the same entry points.  Reference contract: grad_dft/molecule.py:385-409 (jit-wrapped primitives) and
value_and_grad, grad_dft/train.py:86-121."""
import sys
from pathlib import Path

import pytest
import torch

from graddft_b200 import jax_ffi, ops
from graddft_b200._lib import GDFT_GRAD, GDFT_HF, GDFT_LAPL, GDFT_RHO, GDFT_TAU
from graddft_b200.synthetic import synthetic_molecule

sys.path.insert(0, str(Path(__file__).resolve().parent))
import jax_min_shim as shim  # noqa: E402

pytestmark = pytest.mark.gpu
F64 = torch.float64


@pytest.fixture()
def jx(cuda_device, monkeypatch):
    jax, jnp = shim.build(cuda_device)
    monkeypatch.setattr(jax_ffi, "jax", jax)
    monkeypatch.setattr(jax_ffi, "jnp", jnp)
    monkeypatch.setattr(jax_ffi, "HAVE_JAX", True)
    shim.CALLS.clear()
    jax_ffi.register()
    assert len(jax.ffi.registered) == len(jax_ffi._TARGETS) == 24
    return jax


@pytest.fixture(scope="module")
def mol(cuda_device):
    m = synthetic_molecule(900, 21, n_omega=2, seed=1984, device=cuda_device, mask_frac=0.0)
    return m, ops.PackedBasis(m["ao"], m["grad_ao"], m["grad_n_ao2"], m["chi"])


def rnd(shape, dev, seed):
    return torch.randn(shape, dtype=F64, device=dev, generator=torch.Generator(device=dev).manual_seed(seed))


def test_pack_and_density_family(jx, mol, cuda_device):
    m, basis = mol
    jb = jax_ffi.pack_basis(m["ao"], m["grad_ao"], m["grad_n_ao2"], m["chi"])
    assert torch.equal(jb.planes, basis.planes) and torch.equal(jb.chi_packed, basis.chi_packed) and (jb.N, jb.n, jb.nplanes, jb.W) == (900, 21, 5, 2)
    flags = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL | GDFT_HF
    out = jax_ffi.density_family(jb, m["rdm1"], flags)
    ref = ops._density_fwd_raw(basis, m["rdm1"], flags)
    assert all(torch.equal(a, b) for a, b in zip(out, ref))
    sub = jax_ffi.density_family(jb, m["rdm1"], GDFT_RHO | GDFT_GRAD)           # unselected outputs come back as None
    assert torch.equal(sub[0], ref[0]) and torch.equal(sub[1], ref[1]) and sub[2] is None and sub[4] is None
    # first-order rule: the pullback is gdft_density_bwd (+ gdft_hf_fock summed over omega) ...
    jax_ffi.density_family(jb, m["rdm1"], flags)
    f, args = shim.CALLS[-1]
    primal, pull = shim.vjp(f, *args)
    cots = tuple(rnd(t.shape, cuda_device, 10 + i) for i, t in enumerate(ref))
    (dbar,) = pull(cots)
    want = ops._density_bwd_raw(basis, flags & ~GDFT_HF, *cots[:4]) + ops._hf_fock_raw(basis, cots[4]).sum(0)
    assert torch.equal(dbar, want)
    # ... and the rule of THAT call is the forward map again (closure under repeated differentiation)
    g, gargs = shim.CALLS[-1]
    assert g is not f
    _, pull2 = shim.vjp(g, *gargs)
    dd = rnd((2, 21, 21), cuda_device, 3)
    (back,) = pull2(dd)
    again = ops._density_fwd_raw(basis, dd, flags)
    assert all(torch.equal(a, b) for a, b in zip(back, again))


def test_coulomb_and_quadrature_rules(jx, mol, cuda_device):
    m, _ = mol
    eri, P = m["rep_tensor"].contiguous(), m["rdm1"].sum(0).contiguous()
    J = jax_ffi.coulomb_j(eri, P)
    assert torch.equal(J, ops._eri_j_raw(P, eri)[0])
    f, args = shim.CALLS[-1]
    _, pull = shim.vjp(f, *args)
    Jb = rnd(J.shape, cuda_device, 5)
    (Pb,) = pull(Jb)
    assert torch.equal(Pb, ops._eri_jt_raw(Jb, eri))
    # J and K from one sweep: values, the pullback through both outputs, and the rules of the transposed sweeps (closure)
    J2, K2 = jax_ffi.coulomb_jk(eri, P)
    Jk_ref, K_ref = ops._eri_jk_raw(P, eri)
    assert torch.equal(J2, Jk_ref) and torch.equal(K2, K_ref)
    f, args = shim.CALLS[-1]
    _, pull = shim.vjp(f, *args)
    Kb = rnd(J.shape, cuda_device, 15)
    (Pb2,) = pull((Jb, Kb))
    assert torch.equal(Pb2, ops._eri_jt_raw(Jb, eri) + ops._eri_kt_raw(Kb, eri))
    g, gargs = shim.CALLS[-1]  # the K transpose the pullback just called
    assert g is not f
    (back,) = shim.vjp(g, *gargs)[1](P)
    assert torch.equal(back, K_ref)
    rows = 77
    block = eri.reshape(21 * 21, 21, 21)[40:40 + rows].contiguous()
    Jr = jax_ffi.coulomb_j_rows(block, P)
    assert torch.equal(Jr, J.reshape(-1)[40:40 + rows])
    f, args = shim.CALLS[-1]
    (Pr,) = shim.vjp(f, *args)[1](rnd((rows,), cuda_device, 6))
    assert torch.equal(Pr, ops._CoulombJRowsT.apply(rnd((rows,), cuda_device, 6), block))

    N, F = 900, 5
    d, w = rnd((N, F), cuda_device, 7), m["weights"].contiguous()
    for c_rows in (1, N):
        c = rnd((c_rows, F), cuda_device, 8)
        E = jax_ffi.xc_integrate(c, d, w)
        cl, dl = c.clone().requires_grad_(True), d.clone().requires_grad_(True)
        E_ref = ops.xc_integrate(cl, dl, w)
        assert torch.equal(E, E_ref.detach())
        f, args = shim.CALLS[-1]
        cb, db = shim.vjp(f, *args)[1](torch.tensor(1.3, dtype=F64, device=cuda_device))
        cb_ref, db_ref = torch.autograd.grad(E_ref, (cl, dl), torch.tensor(1.3, dtype=F64, device=cuda_device))
        assert torch.equal(cb, cb_ref) and torch.equal(db, db_ref)


def test_pointwise_rules_to_second_order(jx, mol, cuda_device):
    m, basis = mol
    rho, grho, tau, lapl, _ = ops._density_fwd_raw(basis, m["rdm1"], GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL)
    for name, args in (("B3LYP_SET", (rho, grho, None, lapl)), ("DM21_INPUTS", (rho, grho, tau, None)), ("LSDA_X", (rho, None, None, None))):
        leaves = [a.clone().requires_grad_(True) if a is not None else None for a in args]
        ref = ops.pointwise(name, *leaves)
        out = jax_ffi.pointwise(name, *args)
        assert torch.equal(out, ref.detach()), name
        f, fargs = shim.CALLS[-1]
        ob = rnd(out.shape, cuda_device, 21)
        _, pull = shim.vjp(f, *fargs)
        (xbar,) = pull(ob)
        live = [l for l in leaves if l is not None]
        ref_bar = torch.autograd.grad(ref, live, ob, create_graph=True)
        got_bar = [x for x, a in zip(xbar, args) if a is not None]
        assert all(torch.equal(a, b.detach()) for a, b in zip(got_bar, ref_bar)), name
        # second order: the rule of the VJP call (gdft_pointwise_bwd2)
        g, gargs = shim.CALLS[-1]
        assert g is not f
        us = tuple(rnd(x.shape, cuda_device, 30 + i) if a is not None else None for i, (x, a) in enumerate(zip(xbar, args)))
        _, pull2 = shim.vjp(g, *gargs)
        xs_t, ob_t = pull2(us)
        scalar = sum((b * u).sum() for b, u in zip(ref_bar, [u for u in us if u is not None]))
        ref2 = torch.autograd.grad(scalar, live, allow_unused=True)
        got2 = [x for x, a in zip(xs_t, args) if a is not None]
        for a, b in zip(got2, ref2):
            if b is not None:
                assert float((a - b).abs().max()) <= 1e-12 * (1.0 + float(b.abs().max())), name


def test_network_block_and_harness_wrappers(jx, cuda_device):
    N, W = 257, 64
    y, res, scale, bias, ybias = (rnd(s, cuda_device, k) for k, s in enumerate(((N, W), (N, W), (W,), (W,), (W,)), start=40))
    leaves = [t.clone().requires_grad_(True) for t in (y, res, scale, bias, ybias)]
    ref = ops.residual_layernorm_elu(leaves[0], leaves[1], leaves[2], leaves[3], 1e-6, ybias=leaves[4])
    out = jax_ffi.residual_layernorm_elu(y, res, scale, bias, 1e-6, ybias=ybias)
    assert torch.equal(out, ref.detach())
    f, args = shim.CALLS[-1]
    ob = rnd(out.shape, cuda_device, 50)
    zb, ybb, rb, sb, bb = shim.vjp(f, *args)[1](ob)
    gy, gr, gs, gb, gyb = torch.autograd.grad(ref, leaves, ob)
    for a, b in ((zb, gy), (rb, gr), (sb, gs), (bb, gb), (ybb, gyb)):
        assert float((a - b).abs().max()) <= 1e-13 * (1.0 + float(b.abs().max()))
    out2 = jax_ffi.layernorm_elu(y, res, scale, bias, 1e-6)
    assert torch.equal(out2, ops.residual_layernorm_elu(y, res, scale, bias, 1e-6))

    n = 33
    A = rnd((2, n, n), cuda_device, 60)
    A = A + A.transpose(1, 2)
    w_, V_ = jax_ffi.sym_eigh(A)
    wr, Vr = ops.sym_eigh(A)
    assert torch.equal(w_, wr) and torch.equal(V_, Vr)
    mm = 6
    err, fv, x = rnd((mm, 2, n, n), cuda_device, 61), rnd((mm, 2, n, n), cuda_device, 62), rnd((2, mm), cuda_device, 63)
    assert torch.equal(jax_ffi.diis_gram(err), ops.diis_gram(err)) and torch.equal(jax_ffi.diis_combine(x, fv), ops.diis_combine(x, fv))
    ao, D, nu = rnd((40, 20), cuda_device, 64), rnd((2, 20, 20), cuda_device, 65), rnd((40, 20, 20), cuda_device, 66)
    chi_ref = torch.empty(40, 1, 2, 20, dtype=F64, device=cuda_device)
    ops.chi_contract_(chi_ref, 0, 0, ao, D, nu)
    assert torch.equal(jax_ffi.chi_contract(ao, D, nu), chi_ref[:, 0])
