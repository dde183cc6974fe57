"""JAX binding of the C-ABI kernels: XLA custom calls + jax.custom_vjp (the binding BASELINE.json's north-star names;
reference contract: the jit-wrapped primitives of grad_dft/molecule.py:385-409 and `value_and_grad`, grad_dft/train.py:86-121).

Two layers:

1. **Call plans** (`plan_*`, pure Python, no JAX): for every `gdft_*_xla` adapter of csrc/jax_ffi.cu a `Plan` = target name,
   result shapes/dtypes in adapter order, and the packed dims struct (`opaque`).  A plan is everything `jax.ffi.ffi_call`
   needs besides the operands.  Plans are exercised on the GPU WITHOUT JAX by `run_plan_torch` (ctypes, hand-packed
   `void** buffers` exactly as XLA's thunk passes them): tests/test_xla_adapters_gpu.py checks every plan against the direct
   C-ABI entry point, bit for bit.
2. **JAX wrappers** (need `import jax`; JAX is not installed in this repository's build/test environment, SURVEY.md
   section 0.4): `register()` once per process, then the functions below.  Each linear kernel is wrapped with its transpose
   as custom VJP (and vice versa, so `grad(grad(.))` closes over the same two kernels); the per-point maps are bound to
   second order (`pointwise` -> `gdft_pointwise_bwd` -> `gdft_pointwise_bwd2`).  Nothing here contains arithmetic.  The
   wrappers themselves are executed by tests/test_jax_wrappers_gpu.py on a minimal stand-in for `custom_vjp` / `ffi_call`
   (tests/jax_min_shim.py) that routes every call through the real adapters: values, forward/backward rules and their
   nesting are compared with the torch bindings.  What remains unexercised is JAX's own tracing (jit / grad composition).

    from graddft_b200 import jax_ffi
    jax_ffi.register()
    basis = jax_ffi.pack_basis(ao, grad_ao, grad_n_ao[2])            # once per molecule
    rho, grho, tau, lapl = jax_ffi.density_family(basis, rdm1)       # differentiable w.r.t. rdm1, any order
    feats = jax_ffi.pointwise("B3LYP_SET", rho, grho, None, lapl)    # differentiable to second order
    exc = jax_ffi.xc_integrate(coeffs, feats, weights)
"""
from __future__ import annotations

import ctypes
import struct
from typing import NamedTuple, Optional, Sequence, Tuple

from . import _lib
from ._lib import GDFT_GRAD, GDFT_HF, GDFT_LAPL, GDFT_RHO, GDFT_TAU

try:  # pragma: no cover - jax is absent in this environment
    import jax
    import jax.numpy as jnp

    HAVE_JAX = True
except ImportError:  # the only supported state here
    jax = jnp = None
    HAVE_JAX = False

_TARGETS = ("pack_basis", "pack_chi", "density_fwd", "density_bwd", "hf_fock", "eri_j", "eri_j_transpose", "xc_integrate_fwd",
            "xc_integrate_bwd", "pointwise_fwd", "pointwise_bwd", "pointwise_bwd2", "eri_j_rows", "eri_j_transpose_rows", "ln_elu_fwd",
            "ln_elu_bwd", "dense_ln_elu_fwd", "dense_ln_elu_bwd", "sym_eigh", "chi_contract", "diis_gram", "diis_combine", "eri_jk", "eri_k_transpose")
# struct XlaDims { int64 N, n, F, c_rows; int32 flags, nplanes, W, id; double clip; uint64 ws_bytes; }  (jax_ffi.cu)
_DIMS = struct.Struct("<qqqqiiiidQ")
F64, U8 = "float64", "uint8"
DUMMY = ((1,), F64)  # placeholder operand/result of a quantity the flags do not select


def pack_dims(N=0, n=0, F=0, c_rows=0, flags=0, nplanes=0, W=0, pw_id=0, clip=1e-30, ws_bytes=0) -> bytes:
    return _DIMS.pack(N, n, F, c_rows, flags, nplanes, W, pw_id, clip, ws_bytes)


def check_layout() -> None:
    """The packed struct here and in jax_ffi.cu must agree (checked without jax)."""
    assert _lib.lib().gdft_xla_dims_size() == _DIMS.size, (_lib.lib().gdft_xla_dims_size(), _DIMS.size)


class Plan(NamedTuple):
    target: str                                      # gdft_<target>_xla
    n_operands: int
    results: Tuple[Tuple[Tuple[int, ...], str], ...]  # (shape, dtype) in adapter order
    opaque: bytes


def _ws(op: int, N: int, n: int, flags: int, W: int) -> int:
    return max(int(_lib.lib().gdft_workspace_bytes(op, N, n, flags, W)), 256)


def _npad(n: int) -> int:
    return int(_lib.lib().gdft_npad(n))


# ---------------------------------------------------------------------------------------------------------
# call plans (operand order is the adapter's; see the comment above each adapter in csrc/jax_ffi.cu)
# ---------------------------------------------------------------------------------------------------------
def plan_pack_basis(N, n, has_grad=True, has_grad2=True) -> Plan:
    """operands: ao, grad_ao, grad2_ao."""
    nplanes = 1 if not has_grad else (5 if has_grad2 else 4)
    return Plan("pack_basis", 3, (((nplanes, N, _npad(n)), F64),), pack_dims(N=N, n=n, flags=(1 if has_grad else 0) | (2 if has_grad2 else 0), nplanes=nplanes))


def plan_pack_chi(N, n, W) -> Plan:
    """operands: chi[N, W, 2, n]."""
    return Plan("pack_chi", 1, (((W, 2, N, _npad(n)), F64),), pack_dims(N=N, n=n, W=W))


def plan_density_fwd(N, n, nplanes, flags, W=0) -> Plan:
    """operands: packed, rdm1, chi_packed | results: rho, grad_rho, tau, lapl, ehf, ws."""
    ws = _ws(_lib.OP_DENSITY_FWD, N, n, flags, W)
    res = (((N, 2), F64) if flags & GDFT_RHO else DUMMY, ((N, 2, 3), F64) if flags & GDFT_GRAD else DUMMY,
           ((N, 2), F64) if flags & GDFT_TAU else DUMMY, ((N, 2), F64) if flags & GDFT_LAPL else DUMMY,
           ((W, 2, N), F64) if flags & GDFT_HF else DUMMY, ((ws,), U8))
    return Plan("density_fwd", 3, res, pack_dims(N=N, n=n, flags=flags, nplanes=nplanes, W=W if flags & GDFT_HF else 0, ws_bytes=ws))


def plan_density_bwd(N, n, nplanes, flags) -> Plan:
    """operands: packed, rho_bar, grad_rho_bar, tau_bar, lapl_bar | results: rdm1_bar, ws."""
    ws = _ws(_lib.OP_DENSITY_BWD, N, n, flags, 0)
    return Plan("density_bwd", 5, (((2, n, n), F64), ((ws,), U8)), pack_dims(N=N, n=n, flags=flags, nplanes=nplanes, ws_bytes=ws))


def plan_hf_fock(N, n, nplanes, W) -> Plan:
    """operands: packed, chi_packed, g[W, 2, N] | results: fock[W, 2, n, n], ws."""
    ws = _ws(_lib.OP_HF_FOCK, N, n, 0, W)
    return Plan("hf_fock", 3, (((W, 2, n, n), F64), ((ws,), U8)), pack_dims(N=N, n=n, W=W, nplanes=nplanes, ws_bytes=ws))


def plan_eri_j(n) -> Plan:
    """operands: eri[n, n, n, n], P[n, n] | results: J[n, n], E_J[1]."""
    return Plan("eri_j", 2, (((n, n), F64), ((1,), F64)), pack_dims(n=n))


def plan_eri_j_transpose(n) -> Plan:
    """operands: eri, J_bar | results: P_bar, ws."""
    ws = _ws(_lib.OP_ERI_J, 0, n, 0, 0)
    return Plan("eri_j_transpose", 2, (((n, n), F64), ((ws,), U8)), pack_dims(n=n, ws_bytes=ws))


def plan_eri_jk(n) -> Plan:
    """operands: eri[n, n, n, n], P[n, n] | results: J[n, n], K[n, n], ws (one pass over the tensor for both)."""
    ws = _ws(_lib.OP_ERI_J, 0, n, 0, 0)
    return Plan("eri_jk", 2, (((n, n), F64), ((n, n), F64), ((ws,), U8)), pack_dims(n=n, ws_bytes=ws))


def plan_eri_k_transpose(n) -> Plan:
    """operands: eri, K_bar | results: P_bar, ws."""
    ws = _ws(_lib.OP_ERI_J, 0, n, 0, 0)
    return Plan("eri_k_transpose", 2, (((n, n), F64), ((ws,), U8)), pack_dims(n=n, ws_bytes=ws))


def plan_eri_j_rows(n, rows) -> Plan:
    """operands: eri_rows[rows, n, n], P | results: J_rows[rows]."""
    return Plan("eri_j_rows", 2, (((rows,), F64),), pack_dims(N=rows, n=n))


def plan_eri_j_transpose_rows(n, rows) -> Plan:
    """operands: eri_rows, J_bar_rows | results: P_bar, ws."""
    ws = _ws(_lib.OP_ERI_J, 0, n, 0, 0)
    return Plan("eri_j_transpose_rows", 2, (((n, n), F64), ((ws,), U8)), pack_dims(N=rows, n=n, ws_bytes=ws))


def plan_xc_integrate_fwd(N, F, c_rows, clip=1e-30) -> Plan:
    """operands: c[c_rows, F], d[N, F], w[N] | results: E[1], ws."""
    ws = _ws(_lib.OP_XC_INTEGRATE, N, 0, 0, 0)
    return Plan("xc_integrate_fwd", 3, (((1,), F64), ((ws,), U8)), pack_dims(N=N, F=F, c_rows=c_rows, clip=clip, ws_bytes=ws))


def plan_xc_integrate_bwd(N, F, c_rows, clip=1e-30) -> Plan:
    """operands: c, d, w, E_bar[1] | results: c_bar[c_rows, F], d_bar[N, F], ws."""
    ws = _ws(_lib.OP_XC_INTEGRATE, N, 0, 0, 0)
    return Plan("xc_integrate_bwd", 4, (((c_rows, F), F64), ((N, F), F64), ((ws,), U8)), pack_dims(N=N, F=F, c_rows=c_rows, clip=clip, ws_bytes=ws))


def _pw_flags(has_grad, has_tau, has_lapl) -> int:
    return (1 if has_grad else 0) | (2 if has_lapl else 0) | (4 if has_tau else 0)


def _pw_id(name) -> int:
    return _lib.PW_IDS[name] if isinstance(name, str) else int(name)


def plan_pointwise_fwd(name, N, has_grad, has_tau, has_lapl, clip=1e-30) -> Plan:
    """operands: rho, grad_rho, tau, lapl | results: out[N, F]."""
    pid = _pw_id(name)
    F = int(_lib.lib().gdft_pointwise_ncols(pid))
    return Plan("pointwise_fwd", 4, (((N, F), F64),), pack_dims(N=N, pw_id=pid, flags=_pw_flags(has_grad, has_tau, has_lapl), clip=clip))


def _pw_cots(N, has_grad, has_tau, has_lapl):
    return (((N, 2), F64), ((N, 2, 3), F64) if has_grad else DUMMY, ((N, 2), F64) if has_tau else DUMMY, ((N, 2), F64) if has_lapl else DUMMY)


def plan_pointwise_bwd(name, N, has_grad, has_tau, has_lapl, clip=1e-30) -> Plan:
    """operands: rho, grad_rho, tau, lapl, out_bar | results: rho_bar, grad_rho_bar, tau_bar, lapl_bar."""
    return Plan("pointwise_bwd", 5, _pw_cots(N, has_grad, has_tau, has_lapl),
                pack_dims(N=N, pw_id=_pw_id(name), flags=_pw_flags(has_grad, has_tau, has_lapl), clip=clip))


def plan_pointwise_bwd2(name, N, has_grad, has_tau, has_lapl, clip=1e-30) -> Plan:
    """operands: rho, grad_rho, tau, lapl, out_bar, u_rho, u_grad_rho, u_tau, u_lapl | results: out_bar_bar[N, F], rho_t, grad_rho_t,
    tau_t, lapl_t (the VJP of the per-point VJP: J.u and the mixed second derivatives)."""
    pid = _pw_id(name)
    F = int(_lib.lib().gdft_pointwise_ncols(pid))
    return Plan("pointwise_bwd2", 9, (((N, F), F64),) + _pw_cots(N, has_grad, has_tau, has_lapl),
                pack_dims(N=N, pw_id=pid, flags=_pw_flags(has_grad, has_tau, has_lapl), clip=clip))


def plan_ln_elu_fwd(N, W, has_res, eps=1e-6) -> Plan:
    """operands: y, res, scale, bias | results: out[N, W], stats[N, 2]."""
    return Plan("ln_elu_fwd", 4, (((N, W), F64), ((N, 2), F64)), pack_dims(N=N, n=W, flags=1 if has_res else 0, clip=eps))


def plan_ln_elu_bwd(N, W, has_res, eps=1e-6) -> Plan:
    """operands: y, res, scale, bias, stats, out_bar | results: z_bar, scale_bar, bias_bar, ws."""
    ws = _ws(_lib.OP_LN_ELU, N, W, 0, 0)
    return Plan("ln_elu_bwd", 6, (((N, W), F64), ((W,), F64), ((W,), F64), ((ws,), U8)), pack_dims(N=N, n=W, flags=1 if has_res else 0, clip=eps, ws_bytes=ws))


def plan_dense_ln_elu_fwd(N, W, has_res, has_ybias, eps=1e-6) -> Plan:
    """operands: y, ybias, res, scale, bias | results: out, stats."""
    return Plan("dense_ln_elu_fwd", 5, (((N, W), F64), ((N, 2), F64)), pack_dims(N=N, n=W, flags=(1 if has_res else 0) | (2 if has_ybias else 0), clip=eps))


def plan_dense_ln_elu_bwd(N, W, has_res, has_ybias, has_fwd_out, eps=1e-6) -> Plan:
    """operands: y, ybias, res, scale, bias, stats, fwd_out, out_bar | results: z_bar, scale_bar, bias_bar, ybias_bar, ws."""
    ws = _ws(_lib.OP_LN_ELU, N, W, 0, 0)
    flags = (1 if has_res else 0) | (2 if has_ybias else 0) | (4 if has_fwd_out else 0)
    return Plan("dense_ln_elu_bwd", 8, (((N, W), F64), ((W,), F64), ((W,), F64), ((W,), F64) if has_ybias else DUMMY, ((ws,), U8)),
                pack_dims(N=N, n=W, flags=flags, clip=eps, ws_bytes=ws))


def plan_sym_eigh(batch, n) -> Plan:
    """operands: A[batch, n, n] | results: evals[batch, n], evecs[batch, n, n]."""
    return Plan("sym_eigh", 1, (((batch, n), F64), ((batch, n, n), F64)), pack_dims(N=batch, n=n))


def plan_chi_contract(rows, n) -> Plan:
    """operands: ao[rows, n], rdm1[2, n, n], nu[rows, n, n] | results: chi[rows, 2, n] (one omega, one chunk)."""
    return Plan("chi_contract", 3, (((rows, 2, n), F64),), pack_dims(N=rows, n=n))


def plan_diis_gram(m, n) -> Plan:
    """operands: err_vec[m, 2, n, n] | results: gram[2, m, m]."""
    return Plan("diis_gram", 1, (((2, m, m), F64),), pack_dims(W=m, n=n))


def plan_diis_combine(m, n) -> Plan:
    """operands: x[2, m], fock_vec[m, 2, n, n] | results: out[2, n, n]."""
    return Plan("diis_combine", 2, (((2, n, n), F64),), pack_dims(W=m, n=n))


def run_plan_torch(plan: Plan, operands: Sequence) -> list:
    """Execute a plan through the XLA adapter with torch tensors and ctypes (the test driver; what XLA's thunk does)."""
    import torch

    L = _lib.lib()
    assert len(operands) == plan.n_operands, (plan.target, len(operands), plan.n_operands)
    dev = operands[0].device
    results = [torch.empty(shape, dtype=getattr(torch, dt), device=dev) for shape, dt in plan.results]
    keep = [o.contiguous() for o in operands]
    bufs = keep + results
    arr = (ctypes.c_void_p * len(bufs))(*[ctypes.c_void_p(t.data_ptr()) for t in bufs])
    getattr(L, f"gdft_{plan.target}_xla")(_lib.stream_ptr(), arr, plan.opaque, len(plan.opaque))
    status = L.gdft_xla_last_status()
    if status != 0:
        raise _lib.GdftError(f"gdft_{plan.target}_xla: {L.gdft_status_string(status).decode()}")
    return results


# ---------------------------------------------------------------------------------------------------------
# JAX side
# ---------------------------------------------------------------------------------------------------------
def register() -> None:  # pragma: no cover
    if not HAVE_JAX:
        raise ImportError("jax is not installed; use the torch bindings in graddft_b200.ops")
    check_layout()
    L = _lib.lib()
    for name in _TARGETS:
        fn = getattr(L, f"gdft_{name}_xla")
        capsule = jax.ffi.pycapsule(ctypes.cast(fn, ctypes.c_void_p).value) if hasattr(jax.ffi, "pycapsule") else fn
        jax.ffi.register_ffi_target(f"gdft_{name}", capsule, platform="CUDA", api_version=0)


def _run(plan: Plan, *operands):  # pragma: no cover
    shapes = tuple(jax.ShapeDtypeStruct(s, jnp.dtype(dt)) for s, dt in plan.results)
    assert len(operands) == plan.n_operands
    return jax.ffi.ffi_call(f"gdft_{plan.target}", shapes, custom_call_api_version=2, legacy_backend_config=plan.opaque)(*operands)


def _dummy():  # pragma: no cover
    return jnp.zeros((1,), jnp.float64)


def _or_dummy(x):  # pragma: no cover
    return _dummy() if x is None else x


class PackedBasis(NamedTuple):  # pragma: no cover
    """Device-resident packed planes of one molecule (the JAX twin of ops.PackedBasis)."""
    planes: object
    chi_packed: Optional[object]
    N: int
    n: int
    nplanes: int
    W: int


def pack_basis(ao, grad_ao=None, grad2_ao=None, chi=None) -> "PackedBasis":  # pragma: no cover
    N, n = ao.shape
    p = plan_pack_basis(N, n, grad_ao is not None, grad2_ao is not None)
    (planes,) = _run(p, ao, _or_dummy(grad_ao), _or_dummy(grad2_ao))
    chi_packed, W = None, 0
    if chi is not None:
        W = chi.shape[1]
        (chi_packed,) = _run(plan_pack_chi(N, n, W), chi)
    return PackedBasis(planes, chi_packed, N, n, p.results[0][0][0], W)


def _density_flags(basis, flags):  # pragma: no cover
    if flags is None:
        flags = GDFT_RHO | (GDFT_GRAD | GDFT_TAU if basis.nplanes >= 4 else 0) | (GDFT_LAPL if basis.nplanes >= 5 else 0)
    return flags


def density_family(basis, rdm1, flags: Optional[int] = None):  # pragma: no cover
    """(rho, grad_rho, tau, lapl, e_HF) selected by `flags` (None for unselected) -- grad_dft/molecule.py:409,440,502,
    472-474,537-541 in ONE launch.  Linear in rdm1; its VJP is gdft_density_bwd (+ gdft_hf_fock for e_HF), whose VJP is this call."""
    flags = _density_flags(basis, flags)
    N, n = basis.N, basis.n
    sel = [bool(flags & f) for f in (GDFT_RHO, GDFT_GRAD, GDFT_TAU, GDFT_LAPL, GDFT_HF)]

    def raw_fwd(D):
        outs = _run(plan_density_fwd(N, n, basis.nplanes, flags, basis.W), basis.planes, D, basis.chi_packed if sel[4] else _dummy())
        return tuple(o if s else None for o, s in zip(outs[:5], sel))

    def raw_bwd(cots):
        fb = flags & ~GDFT_HF
        dbar = jnp.zeros((2, n, n), jnp.float64)
        if fb:
            dbar = _run(plan_density_bwd(N, n, basis.nplanes, fb), basis.planes, *[_or_dummy(c) if s else _dummy() for c, s in zip(cots[:4], sel[:4])])[0]
        if sel[4] and cots[4] is not None:
            # e_HF[w,s,r] = -1/2 sum_ac chi[r,w,s,c] D[s,a,c] ao[r,a]: its transpose is the HF Fock contraction summed over omega
            dbar = dbar + _run(plan_hf_fock(N, n, basis.nplanes, basis.W), basis.planes, basis.chi_packed, cots[4])[0].sum(axis=0)
        return dbar

    @jax.custom_vjp
    def fwd_op(D):
        return raw_fwd(D)

    @jax.custom_vjp
    def bwd_op(cots):
        return raw_bwd(cots)

    fwd_op.defvjp(lambda D: (raw_fwd(D), None), lambda _, cots: (bwd_op(cots),))
    bwd_op.defvjp(lambda cots: (raw_bwd(cots), None), lambda _, dd: (fwd_op(dd),))
    return fwd_op(rdm1)


def density_and_grad(packed, rdm1, N, n, nplanes):  # pragma: no cover
    """(rho[N,2], grad_rho[N,2,3]); kept for callers of the first version of this module."""
    out = density_family(PackedBasis(packed, None, N, n, nplanes, 0), rdm1, GDFT_RHO | GDFT_GRAD)
    return out[0], out[1]


def hf_fock(basis, g):  # pragma: no cover
    """F[w,s,a,c] = -1/2 sum_r ao[r,a] g[w,s,r] chi[r,w,s,c] (grad_dft/molecule.py:606-613, 678-685); linear in g."""
    return _run(plan_hf_fock(basis.N, basis.n, basis.nplanes, basis.W), basis.planes, basis.chi_packed, g)[0]


def coulomb_j(rep_tensor, P):  # pragma: no cover
    """J_pq = sum_rt (pq|rt) P_rt (grad_dft/molecule.py:788-811); the VJP w.r.t. P is the transposed sweep (and vice versa),
    rep_tensor is a constant.  E_J = 1/2 <P, J> (molecule.py:763-783) is left to the caller: 2 n^2 FLOP in jnp."""
    n = P.shape[-1]

    @jax.custom_vjp
    def jt(Jb):
        return _run(plan_eri_j_transpose(n), rep_tensor, Jb)[0]

    @jax.custom_vjp
    def j(Pm):
        return _run(plan_eri_j(n), rep_tensor, Pm)[0]

    j.defvjp(lambda Pm: (j(Pm), None), lambda _, Jb: (jt(Jb),))
    jt.defvjp(lambda Jb: (jt(Jb), None), lambda _, pb: (j(pb),))
    return j(P)


def coulomb_jk(rep_tensor, P):  # pragma: no cover
    """(J, K): J_pq = sum_rt (pq|rt) P_rt and K_pr = sum_qt (pq|rt) P_qt from one pass over rep_tensor (BASELINE.json's "J/K"; the
    reference never forms K from rep_tensor).  Both are linear in P: the VJP is the sum of the two transposed sweeps."""
    n = P.shape[-1]

    @jax.custom_vjp
    def jt(Jb):
        return _run(plan_eri_j_transpose(n), rep_tensor, Jb)[0]

    @jax.custom_vjp
    def kt(Kb):
        return _run(plan_eri_k_transpose(n), rep_tensor, Kb)[0]

    @jax.custom_vjp
    def jk(Pm):
        return tuple(_run(plan_eri_jk(n), rep_tensor, Pm)[:2])

    jk.defvjp(lambda Pm: (jk(Pm), None), lambda _, bars: (jt(bars[0]) + kt(bars[1]),))
    jt.defvjp(lambda Jb: (jt(Jb), None), lambda _, pb: (jk(pb)[0],))
    kt.defvjp(lambda Kb: (kt(Kb), None), lambda _, pb: (jk(pb)[1],))
    return jk(P)


def coulomb_j_rows(rep_rows, P):  # pragma: no cover
    """Row block of J for a (pq)-sharded rep_tensor; VJP = gdft_eri_j_transpose_rows."""
    rows, n = rep_rows.shape[0], P.shape[-1]

    @jax.custom_vjp
    def f(Pm):
        return _run(plan_eri_j_rows(n, rows), rep_rows, Pm)[0]

    f.defvjp(lambda Pm: (f(Pm), None), lambda _, jb: (_run(plan_eri_j_transpose_rows(n, rows), rep_rows, jb)[0],))
    return f(P)


def xc_integrate(c, d, w, clip: float = 1e-30):  # pragma: no cover
    """E = sum_r clip(w_r) clip(sum_f c[r,f] d[r,f]) -- grad_dft/functional.py:219-253, 316-342; VJP w.r.t. c and d."""
    N, F = d.shape
    c_rows = c.shape[0]

    @jax.custom_vjp
    def f(cc, dd):
        return _run(plan_xc_integrate_fwd(N, F, c_rows, clip), cc, dd, w)[0][0]

    def bwd(res, eb):
        cc, dd = res
        cb, db, _ = _run(plan_xc_integrate_bwd(N, F, c_rows, clip), cc, dd, w, jnp.reshape(eb, (1,)))
        return cb, db

    f.defvjp(lambda cc, dd: (f(cc, dd), (cc, dd)), bwd)
    return f(c, d)


def pointwise(name, rho, grad_rho=None, tau=None, lapl=None, clip: float = 1e-30):  # pragma: no cover
    """Closed-form feature set `name` (see _lib.PW_IDS: LSDA_X, B88_X, VWN_C, LYP_C, PW92_C, B3LYP_SET, DM21_INPUTS, DM21_*,
    FEAT_*; grad_dft/popular_functionals.py:29-269, functional.py:504-626, 1048-1202), differentiable to second order."""
    N = rho.shape[0]
    has = (grad_rho is not None, tau is not None, lapl is not None)
    ins = (rho, _or_dummy(grad_rho), _or_dummy(tau), _or_dummy(lapl))

    def pick(outs):
        return tuple(o if (i == 0 or has[i - 1]) else None for i, o in enumerate(outs))

    @jax.custom_vjp
    def vjp_op(xs, ob):
        return tuple(_run(plan_pointwise_bwd(name, N, *has, clip), *xs, ob))

    def vjp_fwd(xs, ob):
        return vjp_op(xs, ob), (xs, ob)

    def vjp_bwd(res, us):
        xs, ob = res
        outs = _run(plan_pointwise_bwd2(name, N, *has, clip), *xs, ob, *[_or_dummy(u) for u in us])
        return tuple(outs[1:5]), outs[0]

    vjp_op.defvjp(vjp_fwd, vjp_bwd)

    @jax.custom_vjp
    def f(xs):
        return _run(plan_pointwise_fwd(name, N, *has, clip), *xs)[0]

    f.defvjp(lambda xs: (f(xs), xs), lambda xs, ob: (vjp_op(xs, ob),))
    return f(ins)


def residual_layernorm_elu(y, res, scale, bias, eps: float = 1e-6, ybias=None):  # pragma: no cover
    """elu(LayerNorm(y [+ ybias] [+ res])) -- the loop body of grad_dft/functional.py:809-819 after the Dense GEMM; first-order VJP."""
    N, W = y.shape
    has_res, has_yb = res is not None, ybias is not None

    @jax.custom_vjp
    def f(yy, yb, rr, sc, bi):
        return _run(plan_dense_ln_elu_fwd(N, W, has_res, has_yb, eps), yy, yb, rr, sc, bi)[0]

    def fwd(yy, yb, rr, sc, bi):
        out, stats = _run(plan_dense_ln_elu_fwd(N, W, has_res, has_yb, eps), yy, yb, rr, sc, bi)
        return out, (yy, yb, rr, sc, bi, stats, out)

    def bwd(resid, ob):
        yy, yb, rr, sc, bi, stats, out = resid
        zb, sb, bb, ybb, _ = _run(plan_dense_ln_elu_bwd(N, W, has_res, has_yb, True, eps), yy, yb, rr, sc, bi, stats, out, ob)
        return zb, (ybb if has_yb else jnp.zeros_like(yb)), (zb if has_res else jnp.zeros_like(rr)), sb, bb

    f.defvjp(fwd, bwd)
    return f(y, _or_dummy(ybias), _or_dummy(res), scale, bias)


def layernorm_elu(y, res, scale, bias, eps: float = 1e-6):  # pragma: no cover
    """The bias-free variant (gdft_ln_elu_fwd / _bwd)."""
    N, W = y.shape
    has_res = res is not None

    @jax.custom_vjp
    def f(yy, rr, sc, bi):
        return _run(plan_ln_elu_fwd(N, W, has_res, eps), yy, rr, sc, bi)[0]

    def fwd(yy, rr, sc, bi):
        out, stats = _run(plan_ln_elu_fwd(N, W, has_res, eps), yy, rr, sc, bi)
        return out, (yy, rr, sc, bi, stats)

    def bwd(resid, ob):
        yy, rr, sc, bi, stats = resid
        zb, sb, bb, _ = _run(plan_ln_elu_bwd(N, W, has_res, eps), yy, rr, sc, bi, stats, ob)
        return zb, (zb if has_res else jnp.zeros_like(rr)), sb, bb

    f.defvjp(fwd, bwd)
    return f(y, _or_dummy(res), scale, bias)


def sym_eigh(A):  # pragma: no cover
    """(evals, evecs) of a batch of symmetric matrices, n <= gdft_sym_eigh_max_n(): the `jnp.linalg.eigh` inside
    grad_dft/utils/eigenproblem.py:26-149, whose own custom VJP (lines 110-149) stays the caller's."""
    batch, n = A.shape[0], A.shape[-1]
    return tuple(_run(plan_sym_eigh(batch, n), A))


def chi_contract(ao_chunk, rdm1, nu_chunk):  # pragma: no cover
    """chi[r, s, a] = sum_bd rdm1[s,b,d] ao[r,b] nu[r,d,a] for one nu chunk and one omega (grad_dft/interface/pyscf.py:1110-1124)."""
    rows, n = ao_chunk.shape
    return _run(plan_chi_contract(rows, n), ao_chunk, rdm1, nu_chunk)[0]


def diis_gram(err_vec):  # pragma: no cover
    """Per-spin Gram matrix of the DIIS error vectors (grad_dft/evaluate.py:1146-1160)."""
    m, n = err_vec.shape[0], err_vec.shape[-1]
    return _run(plan_diis_gram(m, n), err_vec)[0]


def diis_combine(x, fock_vec):  # pragma: no cover
    """sum_i x[s, i] fock_vec[i, s] (grad_dft/evaluate.py:1186-1194)."""
    m, n = fock_vec.shape[0], fock_vec.shape[-1]
    return _run(plan_diis_combine(m, n), x, fock_vec)[0]
