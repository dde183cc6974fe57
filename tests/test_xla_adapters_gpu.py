"""Every XLA custom-call adapter of csrc/jax_ffi.cu (the JAX-FFI boundary BASELINE.json's north-star names, SURVEY.md
section 8b) exercised WITHOUT JAX: the legacy custom-call ABI is `void(stream, void** buffers, const char* opaque, size_t
len)`, so a ctypes caller can hand-pack the buffer list (operands, then results) and the dims struct exactly as XLA's thunk
would.  Each adapter must give bit-identical results to the direct C-ABI entry point it forwards to, leave status 0, and
poison its first result with NaN when the entry point refuses the call (the ABI has no error return).
Reference contract: the jit-wrapped primitives grad_dft/molecule.py:385-409 and value_and_grad, grad_dft/train.py:86-121."""
import ctypes

import pytest
import torch

from graddft_b200 import _lib, jax_ffi, ops
from graddft_b200._lib import GDFT_GRAD, GDFT_HF, GDFT_LAPL, GDFT_RHO, GDFT_TAU
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu
F64 = torch.float64


def xla_call(name, operands, results, **dims):
    L = _lib.lib()
    bufs = list(operands) + list(results)
    arr = (ctypes.c_void_p * len(bufs))(*[ctypes.c_void_p(t.data_ptr()) for t in bufs])
    opaque = jax_ffi.pack_dims(**dims)
    getattr(L, f"gdft_{name}_xla")(_lib.stream_ptr(), arr, opaque, len(opaque))
    torch.cuda.synchronize()
    return L.gdft_xla_last_status()


def dummy(dev):
    return torch.zeros(1, dtype=F64, device=dev)


def ws_tensor(op, N, n, flags, W, dev):
    nbytes = int(_lib.lib().gdft_workspace_bytes(op, N, n, flags, W))
    return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=dev), max(nbytes, 256)


@pytest.fixture(scope="module")
def mol(cuda_device):
    m = synthetic_molecule(700, 19, n_omega=2, seed=1984, device=cuda_device, mask_frac=0.0)
    basis = ops.PackedBasis(m["ao"], m["grad_ao"], m["grad_n_ao2"], m["chi"])
    return m, basis


def test_density_family_adapters(cuda_device, mol):
    m, basis = mol
    dev = cuda_device
    N, n, W = basis.N, basis.n, basis.W
    flags = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL | GDFT_HF
    ref = ops._density_fwd_raw(basis, m["rdm1"], flags)
    outs = [torch.empty_like(t) for t in ref]
    ws, nb = ws_tensor(_lib.OP_DENSITY_FWD, N, n, flags, W, dev)
    assert xla_call("density_fwd", [basis.planes, m["rdm1"].contiguous(), basis.chi_packed], outs + [ws],
                    N=N, n=n, flags=flags, nplanes=basis.nplanes, W=W, ws_bytes=nb) == 0
    for a, b in zip(outs, ref):
        assert torch.equal(a, b)
    # subset (rho + grad only): the unused operands/results are 1-element dummies
    f2 = GDFT_RHO | GDFT_GRAD
    ref2 = ops._density_fwd_raw(basis, m["rdm1"], f2)
    rho, grho = torch.empty_like(ref2[0]), torch.empty_like(ref2[1])
    assert xla_call("density_fwd", [basis.planes, m["rdm1"].contiguous(), dummy(dev)], [rho, grho, dummy(dev), dummy(dev), dummy(dev), ws],
                    N=N, n=n, flags=f2, nplanes=basis.nplanes, W=0, ws_bytes=nb) == 0
    assert torch.equal(rho, ref2[0]) and torch.equal(grho, ref2[1])

    g = torch.Generator(device=dev).manual_seed(7)
    cot = [torch.randn(t.shape, generator=g, dtype=F64, device=dev) for t in ref[:4]]
    fb = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL
    dref = ops._density_bwd_raw(basis, fb, *cot)
    dbar = torch.empty_like(dref)
    ws, nb = ws_tensor(_lib.OP_DENSITY_BWD, N, n, fb, 0, dev)
    assert xla_call("density_bwd", [basis.planes] + cot, [dbar, ws], N=N, n=n, flags=fb, nplanes=basis.nplanes, ws_bytes=nb) == 0
    assert torch.equal(dbar, dref)

    gg = torch.randn((W, 2, N), generator=g, dtype=F64, device=dev)
    fref = ops._hf_fock_raw(basis, gg)
    fock = torch.empty_like(fref)
    ws, nb = ws_tensor(_lib.OP_HF_FOCK, N, n, 0, W, dev)
    assert xla_call("hf_fock", [basis.planes, basis.chi_packed, gg], [fock, ws], N=N, n=n, W=W, nplanes=basis.nplanes, ws_bytes=nb) == 0
    assert torch.equal(fock, fref)


def test_eri_adapters(cuda_device, mol):
    m, _ = mol
    dev = cuda_device
    n = m["rdm1"].shape[-1]
    P = m["rdm1"].sum(0).contiguous()
    eri = m["rep_tensor"].contiguous()
    Jref, EJref = ops._eri_j_raw(P, eri, want_energy=True)
    J, EJ = torch.empty_like(Jref), torch.empty_like(EJref)
    assert xla_call("eri_j", [eri, P], [J, EJ], n=n) == 0
    assert torch.equal(J, Jref) and torch.equal(EJ, EJref)
    Jbar = torch.randn(n, n, dtype=F64, device=dev)
    Pref = ops._eri_jt_raw(Jbar, eri)
    Pbar = torch.empty_like(Pref)
    ws, nb = ws_tensor(_lib.OP_ERI_J, 0, n, 0, 0, dev)
    assert xla_call("eri_j_transpose", [eri, Jbar], [Pbar, ws], n=n, ws_bytes=nb) == 0
    assert torch.equal(Pbar, Pref)
    Jr2, Kr2 = ops._eri_jk_raw(P, eri)
    J2, K2 = torch.empty_like(Jr2), torch.empty_like(Kr2)
    assert xla_call("eri_jk", [eri, P], [J2, K2, ws], n=n, ws_bytes=nb) == 0
    assert torch.equal(J2, Jr2) and torch.equal(K2, Kr2)
    Pk = torch.empty_like(Pref)
    assert xla_call("eri_k_transpose", [eri, Jbar], [Pk, ws], n=n, ws_bytes=nb) == 0
    assert torch.equal(Pk, ops._eri_kt_raw(Jbar, eri))
    rows = 97
    block = eri.reshape(n * n, n, n)[32:32 + rows].contiguous()
    Jr_ref = ops._CoulombJRows.apply(P, block)
    Jr = torch.empty_like(Jr_ref)
    assert xla_call("eri_j_rows", [block, P], [Jr], N=rows, n=n) == 0
    assert torch.equal(Jr, Jr_ref) and torch.equal(Jr, Jref.reshape(-1)[32:32 + rows])
    Jb = torch.randn(rows, dtype=F64, device=dev)
    Pr_ref = ops._CoulombJRowsT.apply(Jb, block)
    Pr = torch.empty_like(Pr_ref)
    assert xla_call("eri_j_transpose_rows", [block, Jb], [Pr, ws], N=rows, n=n, ws_bytes=nb) == 0
    assert torch.equal(Pr, Pr_ref)


def test_quadrature_and_pointwise_adapters(cuda_device, mol):
    m, basis = mol
    dev = cuda_device
    L = _lib.lib()
    N = basis.N
    rho, grho, tau, lapl, _ = ops._density_fwd_raw(basis, m["rdm1"], GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL)
    g = torch.Generator(device=dev).manual_seed(11)
    # K5: B3LYP set needs (rho, grad, lapl): flags bit 0 grad, bit 1 lapl, bit 2 tau; DM21 inputs need (rho, grad, tau)
    for name, fl, args in (("B3LYP_SET", 1 | 2, (rho, grho, None, lapl)), ("DM21_INPUTS", 1 | 4, (rho, grho, tau, None)), ("LSDA_X", 0, (rho, None, None, None))):
        pid = _lib.PW_IDS[name]
        ref = ops.pointwise(name, *args)
        out = torch.empty_like(ref)
        opnds = [a if a is not None else dummy(dev) for a in args]
        assert xla_call("pointwise_fwd", opnds, [out], N=N, pw_id=pid, flags=fl, clip=1e-30) == 0
        assert torch.equal(out, ref), name
        ob = torch.randn(ref.shape, generator=g, dtype=F64, device=dev)
        refb = ops._PointwiseVJP.apply(pid, 1e-30, *args, ob)
        res = [torch.empty_like(t) if t is not None else dummy(dev) for t in refb]  # (rho_bar, grho_bar, tau_bar, lapl_bar)
        assert xla_call("pointwise_bwd", opnds + [ob], res, N=N, pw_id=pid, flags=fl, clip=1e-30) == 0
        for a, b in zip(res, refb):
            if b is not None:
                assert torch.equal(a, b), name
        # second order: cotangents u_* of the VJP's outputs
        us = [torch.randn(t.shape, generator=g, dtype=F64, device=dev) if t is not None else None for t in refb]
        ref2 = [torch.empty_like(ob)] + [torch.empty_like(t) if t is not None else None for t in args]
        assert L.gdft_pointwise_bwd2(_lib.stream_ptr(), N, pid, 1e-30, *[_lib.ptr(a) for a in args], _lib.ptr(ob), *[_lib.ptr(u) for u in us],
                                     *[_lib.ptr(t) for t in ref2]) == 0
        res2 = [torch.empty_like(t) if t is not None else dummy(dev) for t in ref2]
        assert xla_call("pointwise_bwd2", opnds + [ob] + [u if u is not None else dummy(dev) for u in us], res2,
                        N=N, pw_id=pid, flags=fl, clip=1e-30) == 0
        for a, b in zip(res2, ref2):
            if b is not None:
                assert torch.equal(a, b), name

    # K6
    F = 5
    d = torch.randn(N, F, generator=g, dtype=F64, device=dev)
    w = m["weights"].contiguous()
    for c_rows in (1, N):
        c = torch.randn(c_rows, F, generator=g, dtype=F64, device=dev)
        Eref = ops.xc_integrate(c, d, w)
        E = torch.empty(1, dtype=F64, device=dev)
        ws, nb = ws_tensor(_lib.OP_XC_INTEGRATE, N, 0, 0, 0, dev)
        assert xla_call("xc_integrate_fwd", [c, d, w], [E, ws], N=N, F=F, c_rows=c_rows, clip=1e-30, ws_bytes=nb) == 0
        assert torch.equal(E[0], Eref)
        Eb = torch.tensor([0.7], dtype=F64, device=dev)
        cb_ref, db_ref = torch.empty_like(c), torch.empty_like(d)
        assert L.gdft_xc_integrate_bwd(_lib.stream_ptr(), N, F, c_rows, _lib.ptr(c), _lib.ptr(d), _lib.ptr(w), 1e-30, _lib.ptr(Eb), _lib.ptr(cb_ref),
                                       _lib.ptr(db_ref), _lib.wptr(ws), nb) == 0
        cb, db = torch.empty_like(c), torch.empty_like(d)
        assert xla_call("xc_integrate_bwd", [c, d, w, Eb], [cb, db, ws], N=N, F=F, c_rows=c_rows, clip=1e-30, ws_bytes=nb) == 0
        assert torch.equal(cb, cb_ref) and torch.equal(db, db_ref)


def test_network_block_adapters(cuda_device):
    dev = cuda_device
    L = _lib.lib()
    N, W = 333, 64
    g = torch.Generator(device=dev).manual_seed(5)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64, device=dev)  # noqa: E731
    y, res, scale, bias, ybias, ob = rn(N, W), rn(N, W), rn(W), rn(W), rn(W), rn(N, W)
    ws, nb = ws_tensor(_lib.OP_LN_ELU, N, W, 0, 0, dev)
    # plain block
    out_ref, st_ref = torch.empty_like(y), torch.empty(N, 2, dtype=F64, device=dev)
    assert L.gdft_ln_elu_fwd(_lib.stream_ptr(), N, W, _lib.ptr(y), _lib.ptr(res), _lib.ptr(scale), _lib.ptr(bias), 1e-6, _lib.ptr(out_ref), _lib.ptr(st_ref)) == 0
    out, st = torch.empty_like(y), torch.empty_like(st_ref)
    assert xla_call("ln_elu_fwd", [y, res, scale, bias], [out, st], N=N, n=W, flags=1, clip=1e-6) == 0
    assert torch.equal(out, out_ref) and torch.equal(st, st_ref)
    zr, sr, br = torch.empty_like(y), torch.empty_like(scale), torch.empty_like(bias)
    assert L.gdft_ln_elu_bwd(_lib.stream_ptr(), N, W, _lib.ptr(y), _lib.ptr(res), _lib.ptr(scale), _lib.ptr(bias), _lib.ptr(st_ref), _lib.ptr(ob),
                             _lib.ptr(zr), _lib.ptr(sr), _lib.ptr(br), _lib.wptr(ws), nb) == 0
    z, s_, b_ = torch.empty_like(y), torch.empty_like(scale), torch.empty_like(bias)
    assert xla_call("ln_elu_bwd", [y, res, scale, bias, st_ref, ob], [z, s_, b_, ws], N=N, n=W, flags=1, ws_bytes=nb) == 0
    assert torch.equal(z, zr) and torch.equal(s_, sr) and torch.equal(b_, br)
    # with the Dense bias folded in
    assert L.gdft_dense_ln_elu_fwd(_lib.stream_ptr(), N, W, _lib.ptr(y), _lib.ptr(ybias), _lib.ptr(res), _lib.ptr(scale), _lib.ptr(bias), 1e-6,
                                   _lib.ptr(out_ref), _lib.ptr(st_ref)) == 0
    assert xla_call("dense_ln_elu_fwd", [y, ybias, res, scale, bias], [out, st], N=N, n=W, flags=3, clip=1e-6) == 0
    assert torch.equal(out, out_ref) and torch.equal(st, st_ref)
    yr = torch.empty_like(ybias)
    assert L.gdft_dense_ln_elu_bwd(_lib.stream_ptr(), N, W, _lib.ptr(y), _lib.ptr(ybias), _lib.ptr(res), _lib.ptr(scale), _lib.ptr(bias), _lib.ptr(st_ref),
                                   _lib.ptr(out_ref), _lib.ptr(ob), _lib.ptr(zr), _lib.ptr(sr), _lib.ptr(br), _lib.ptr(yr), _lib.wptr(ws), nb) == 0
    yb = torch.empty_like(ybias)
    assert xla_call("dense_ln_elu_bwd", [y, ybias, res, scale, bias, st_ref, out_ref, ob], [z, s_, b_, yb, ws], N=N, n=W, flags=7, ws_bytes=nb) == 0
    assert torch.equal(z, zr) and torch.equal(s_, sr) and torch.equal(b_, br) and torch.equal(yb, yr)


def test_scf_harness_and_chi_adapters(cuda_device):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(3)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64, device=dev)  # noqa: E731
    n = 37
    A = rn(2, n, n)
    A = A + A.transpose(1, 2)
    w_ref, V_ref = ops.sym_eigh(A)
    w, V = torch.empty_like(w_ref), torch.empty_like(V_ref)
    assert xla_call("sym_eigh", [A], [w, V], N=2, n=n) == 0
    assert torch.equal(w, w_ref) and torch.equal(V, V_ref)
    m = 10
    err, fv, x = rn(m, 2, n, n), rn(m, 2, n, n), rn(2, m)
    gram = torch.empty(2, m, m, dtype=F64, device=dev)
    assert xla_call("diis_gram", [err], [gram], W=m, n=n) == 0
    assert torch.equal(gram, ops.diis_gram(err))
    out = torch.empty(2, n, n, dtype=F64, device=dev)
    assert xla_call("diis_combine", [x, fv], [out], W=m, n=n) == 0
    assert torch.equal(out, ops.diis_combine(x, fv))
    Nc, nn = 53, 24
    ao, D, nu = rn(Nc, nn), rn(2, nn, nn), rn(Nc, nn, nn)
    chi_ref = torch.empty(Nc, 1, 2, nn, dtype=F64, device=dev)
    ops.chi_contract_(chi_ref, 0, 0, ao, D, nu)
    chi = torch.empty(Nc, 2, nn, dtype=F64, device=dev)
    assert xla_call("chi_contract", [ao, D, nu], [chi], N=Nc, n=nn) == 0
    assert torch.equal(chi, chi_ref[:, 0])


def test_adapter_error_paths(cuda_device, mol):
    """No error return in this ABI: a refused call poisons its first result with NaN and records the status; a dims struct
    of the wrong size is refused before any buffer is touched."""
    m, basis = mol
    dev = cuda_device
    L = _lib.lib()
    n = basis.n
    P = m["rdm1"].sum(0).contiguous()
    J = torch.zeros(n, n, dtype=F64, device=dev)
    EJ = torch.zeros(1, dtype=F64, device=dev)
    assert xla_call("eri_j", [m["rep_tensor"].contiguous(), P], [J, EJ], n=0) != 0  # bad shape
    assert bool(torch.isnan(J.reshape(-1)[0]))
    arr = (ctypes.c_void_p * 4)(*[ctypes.c_void_p(t.data_ptr()) for t in (m["rep_tensor"], P, J, EJ)])
    J.zero_()
    L.gdft_eri_j_xla(_lib.stream_ptr(), arr, b"short", 5)
    torch.cuda.synchronize()
    assert L.gdft_xla_last_status() == 5 and float(J.abs().max()) == 0.0
    jax_ffi.check_layout()


def test_every_call_plan_of_the_jax_binding(cuda_device, mol):
    """graddft_b200/jax_ffi.py hands `jax.ffi.ffi_call` a Plan (target, result shapes, opaque dims) per adapter.  The very
    same plans executed here through ctypes (`run_plan_torch`) must reproduce the direct entry points bit for bit: that pins
    operand order, result shapes, flag packing and workspace sizes of every JAX wrapper without JAX."""
    m, basis = mol
    dev = cuda_device
    J = jax_ffi
    run = J.run_plan_torch
    N, n, W = basis.N, basis.n, basis.W
    d1 = dummy(dev)
    g = torch.Generator(device=dev).manual_seed(23)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64, device=dev)  # noqa: E731
    covered = set()

    def go(plan, ops_):
        covered.add(plan.target)
        out = run(plan, ops_)
        torch.cuda.synchronize()
        return out

    (planes,) = go(J.plan_pack_basis(N, n, True, True), [m["ao"], m["grad_ao"], m["grad_n_ao2"]])
    assert torch.equal(planes, basis.planes)
    (chi_packed,) = go(J.plan_pack_chi(N, n, W), [m["chi"]])
    assert torch.equal(chi_packed, basis.chi_packed)
    flags = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL | GDFT_HF
    ref = ops._density_fwd_raw(basis, m["rdm1"], flags)
    outs = go(J.plan_density_fwd(N, n, basis.nplanes, flags, W), [planes, m["rdm1"], chi_packed])
    assert all(torch.equal(a, b) for a, b in zip(outs[:5], ref))
    outs = go(J.plan_density_fwd(N, n, basis.nplanes, GDFT_RHO, W), [planes, m["rdm1"], d1])
    assert torch.equal(outs[0], ref[0]) and outs[1].numel() == 1
    fb = GDFT_RHO | GDFT_GRAD | GDFT_TAU | GDFT_LAPL
    cot = [rn(*t.shape) for t in ref[:4]]
    assert torch.equal(go(J.plan_density_bwd(N, n, basis.nplanes, fb), [planes] + cot)[0], ops._density_bwd_raw(basis, fb, *cot))
    assert torch.equal(go(J.plan_density_bwd(N, n, basis.nplanes, GDFT_GRAD), [planes, d1, cot[1], d1, d1])[0],
                       ops._density_bwd_raw(basis, GDFT_GRAD, None, cot[1], None, None))
    gg = rn(W, 2, N)
    assert torch.equal(go(J.plan_hf_fock(N, n, basis.nplanes, W), [planes, chi_packed, gg])[0], ops._hf_fock_raw(basis, gg))

    P, eri = m["rdm1"].sum(0).contiguous(), m["rep_tensor"].contiguous()
    Jref, EJref = ops._eri_j_raw(P, eri, want_energy=True)
    Jm, EJ = go(J.plan_eri_j(n), [eri, P])
    assert torch.equal(Jm, Jref) and torch.equal(EJ, EJref)
    Jbar = rn(n, n)
    assert torch.equal(go(J.plan_eri_j_transpose(n), [eri, Jbar])[0], ops._eri_jt_raw(Jbar, eri))
    assert all(torch.equal(x, y) for x, y in zip(go(J.plan_eri_jk(n), [eri, P])[:2], ops._eri_jk_raw(P, eri)))
    assert torch.equal(go(J.plan_eri_k_transpose(n), [eri, Jbar])[0], ops._eri_kt_raw(Jbar, eri))
    rows = 61
    block = eri.reshape(n * n, n, n)[17:17 + rows].contiguous()
    assert torch.equal(go(J.plan_eri_j_rows(n, rows), [block, P])[0], Jref.reshape(-1)[17:17 + rows])
    Jb = rn(rows)
    assert torch.equal(go(J.plan_eri_j_transpose_rows(n, rows), [block, Jb])[0], ops._CoulombJRowsT.apply(Jb, block))

    rho, grho, tau, lapl = ref[:4]
    for name, args in (("B3LYP_SET", (rho, grho, None, lapl)), ("DM21_INPUTS", (rho, grho, tau, None)), ("FEAT_MGGA", (rho, grho, tau, None))):
        has = (args[1] is not None, args[2] is not None, args[3] is not None)
        opnds = [a if a is not None else d1 for a in args]
        want = ops.pointwise(name, *args)
        (out,) = go(J.plan_pointwise_fwd(name, N, *has), opnds)
        assert torch.equal(out, want), name
        ob = rn(*want.shape)
        refb = ops._PointwiseVJP.apply(_lib.PW_IDS[name], 1e-30, *args, ob)
        res = go(J.plan_pointwise_bwd(name, N, *has), opnds + [ob])
        assert all(torch.equal(a, b) for a, b in zip(res, refb) if b is not None), name
        us = [rn(*t.shape) if t is not None else None for t in refb]
        ref2 = [torch.empty_like(ob)] + [torch.empty_like(t) if t is not None else None for t in args]
        assert _lib.lib().gdft_pointwise_bwd2(_lib.stream_ptr(), N, _lib.PW_IDS[name], 1e-30, *[_lib.ptr(a) for a in args], _lib.ptr(ob),
                                              *[_lib.ptr(u) for u in us], *[_lib.ptr(t) for t in ref2]) == 0
        res2 = go(J.plan_pointwise_bwd2(name, N, *has), opnds + [ob] + [u if u is not None else d1 for u in us])
        assert all(torch.equal(a, b) for a, b in zip(res2, ref2) if b is not None), name

    F = 4
    d, w = rn(N, F), m["weights"].contiguous()
    for c_rows in (1, N):
        c = rn(c_rows, F)
        cl, dl = c.clone().requires_grad_(True), d.clone().requires_grad_(True)
        E_ref = ops.xc_integrate(cl, dl, w)
        assert torch.equal(go(J.plan_xc_integrate_fwd(N, F, c_rows), [c, d, w])[0][0], E_ref.detach())
        cb_ref, db_ref = torch.autograd.grad(E_ref, (cl, dl), torch.tensor(0.7, dtype=F64, device=dev))
        cb, db, _ = go(J.plan_xc_integrate_bwd(N, F, c_rows), [c, d, w, torch.tensor([0.7], dtype=F64, device=dev)])
        assert torch.equal(cb, cb_ref) and torch.equal(db, db_ref)

    Nn, Wd = 301, 64
    L = _lib.lib()
    y, res_, scale, bias, ybias, ob = rn(Nn, Wd), rn(Nn, Wd), rn(Wd), rn(Wd), rn(Wd), rn(Nn, Wd)
    out_ref, st_ref = torch.empty_like(y), torch.empty(Nn, 2, dtype=F64, device=dev)
    assert L.gdft_ln_elu_fwd(_lib.stream_ptr(), Nn, Wd, _lib.ptr(y), _lib.ptr(res_), _lib.ptr(scale), _lib.ptr(bias), 1e-6, _lib.ptr(out_ref), _lib.ptr(st_ref)) == 0
    out, st = go(J.plan_ln_elu_fwd(Nn, Wd, True), [y, res_, scale, bias])
    assert torch.equal(out, out_ref) and torch.equal(st, st_ref)
    ws, nb = ws_tensor(_lib.OP_LN_ELU, Nn, Wd, 0, 0, dev)
    zr, sr, br, yr = torch.empty_like(y), torch.empty_like(scale), torch.empty_like(bias), torch.empty_like(ybias)
    assert L.gdft_ln_elu_bwd(_lib.stream_ptr(), Nn, Wd, _lib.ptr(y), _lib.ptr(res_), _lib.ptr(scale), _lib.ptr(bias), _lib.ptr(st_ref), _lib.ptr(ob),
                             _lib.ptr(zr), _lib.ptr(sr), _lib.ptr(br), _lib.wptr(ws), nb) == 0
    z, s_, b_, _ = go(J.plan_ln_elu_bwd(Nn, Wd, True), [y, res_, scale, bias, st_ref, ob])
    assert torch.equal(z, zr) and torch.equal(s_, sr) and torch.equal(b_, br)
    assert L.gdft_dense_ln_elu_fwd(_lib.stream_ptr(), Nn, Wd, _lib.ptr(y), _lib.ptr(ybias), _lib.ptr(res_), _lib.ptr(scale), _lib.ptr(bias), 1e-6,
                                   _lib.ptr(out_ref), _lib.ptr(st_ref)) == 0
    out, st = go(J.plan_dense_ln_elu_fwd(Nn, Wd, True, True), [y, ybias, res_, scale, bias])
    assert torch.equal(out, out_ref) and torch.equal(st, st_ref)
    assert L.gdft_dense_ln_elu_bwd(_lib.stream_ptr(), Nn, Wd, _lib.ptr(y), _lib.ptr(ybias), _lib.ptr(res_), _lib.ptr(scale), _lib.ptr(bias), _lib.ptr(st_ref),
                                   _lib.ptr(out_ref), _lib.ptr(ob), _lib.ptr(zr), _lib.ptr(sr), _lib.ptr(br), _lib.ptr(yr), _lib.wptr(ws), nb) == 0
    z, s_, b_, yb, _ = go(J.plan_dense_ln_elu_bwd(Nn, Wd, True, True, True), [y, ybias, res_, scale, bias, st_ref, out_ref, ob])
    assert torch.equal(z, zr) and torch.equal(s_, sr) and torch.equal(b_, br) and torch.equal(yb, yr)

    ne = 29
    A = rn(2, ne, ne)
    A = A + A.transpose(1, 2)
    w_ref, V_ref = ops.sym_eigh(A)
    wv, V = go(J.plan_sym_eigh(2, ne), [A])
    assert torch.equal(wv, w_ref) and torch.equal(V, V_ref)
    mm = 7
    err, fv, x = rn(mm, 2, ne, ne), rn(mm, 2, ne, ne), rn(2, mm)
    assert torch.equal(go(J.plan_diis_gram(mm, ne), [err])[0], ops.diis_gram(err))
    assert torch.equal(go(J.plan_diis_combine(mm, ne), [x, fv])[0], ops.diis_combine(x, fv))
    Nc, nn = 41, 22
    ao, D, nu = rn(Nc, nn), rn(2, nn, nn), rn(Nc, nn, nn)
    chi_ref = torch.empty(Nc, 1, 2, nn, dtype=F64, device=dev)
    ops.chi_contract_(chi_ref, 0, 0, ao, D, nu)
    assert torch.equal(go(J.plan_chi_contract(Nc, nn), [ao, D, nu])[0], chi_ref[:, 0])
    assert covered == set(J._TARGETS), set(J._TARGETS) - covered
