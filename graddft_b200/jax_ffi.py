"""JAX binding of the C-ABI kernels: XLA custom calls + jax.custom_vjp (the binding BASELINE.json's north-star names).

JAX is NOT installed in this repository's build/test environment, so this module is
importable only where `import jax` works and is not exercised by the test-suite here; the torch bindings in
`ops.py` drive the very same entry points and are the tested path.  It is kept deliberately thin: every function
below only packs dimensions and declares result shapes -- the arithmetic is in libgdft_b200.so.

Usage (in an environment with jax[cuda] >= 0.4.31):

    from graddft_b200 import jax_ffi
    jax_ffi.register()                                   # once per process
    packed = jax_ffi.pack_basis(ao, grad_ao, grad_2_ao)  # once per molecule
    rho, grho = jax_ffi.density_and_grad(packed, rdm1)   # differentiable w.r.t. rdm1 (custom_vjp -> gdft_density_bwd)
"""
from __future__ import annotations

import ctypes
import struct

from . import _lib

try:  # pragma: no cover - jax is absent in this environment
    import jax
    import jax.numpy as jnp

    HAVE_JAX = True
except ImportError:  # the only supported state here
    jax = jnp = None
    HAVE_JAX = False

_TARGETS = ("density_fwd", "density_bwd", "hf_fock", "eri_j", "eri_j_transpose", "xc_integrate_fwd", "xc_integrate_bwd",
            "pointwise_fwd", "pointwise_bwd", "pointwise_bwd2", "eri_j_rows", "eri_j_transpose_rows", "ln_elu_fwd", "ln_elu_bwd",
            "dense_ln_elu_fwd", "dense_ln_elu_bwd", "sym_eigh", "chi_contract", "diis_gram", "diis_combine")
# struct XlaDims { int64 N, n, F, c_rows; int32 flags, nplanes, W, id; double clip; uint64 ws_bytes; }  (jax_ffi.cu)
_DIMS = struct.Struct("<qqqqiiiidQ")


def pack_dims(N=0, n=0, F=0, c_rows=0, flags=0, nplanes=0, W=0, pw_id=0, clip=1e-30, ws_bytes=0) -> bytes:
    return _DIMS.pack(N, n, F, c_rows, flags, nplanes, W, pw_id, clip, ws_bytes)


def check_layout() -> None:
    """The packed struct here and in jax_ffi.cu must agree (checked without jax)."""
    assert _lib.lib().gdft_xla_dims_size() == _DIMS.size, (_lib.lib().gdft_xla_dims_size(), _DIMS.size)


def register() -> None:  # pragma: no cover
    if not HAVE_JAX:
        raise ImportError("jax is not installed; use the torch bindings in graddft_b200.ops")
    check_layout()
    L = _lib.lib()
    for name in _TARGETS:
        fn = getattr(L, f"gdft_{name}_xla")
        capsule = jax.ffi.pycapsule(ctypes.cast(fn, ctypes.c_void_p).value) if hasattr(jax.ffi, "pycapsule") else fn
        jax.ffi.register_ffi_target(f"gdft_{name}", capsule, platform="CUDA", api_version=0)


def _call(name, result_shapes, *operands, opaque: bytes):  # pragma: no cover
    return jax.ffi.ffi_call(f"gdft_{name}", result_shapes, custom_call_api_version=2, legacy_backend_config=opaque)(*operands)


def density_and_grad(packed, rdm1, N, n, nplanes):  # pragma: no cover
    """(rho[N,2], grad_rho[N,2,3]) with a custom VJP that calls gdft_density_bwd (whose own VJP is gdft_density_fwd)."""
    L = _lib.lib()
    flags = _lib.GDFT_RHO | _lib.GDFT_GRAD
    f64 = jnp.float64

    @jax.custom_vjp
    def fwd_op(rdm1):
        ws = int(L.gdft_workspace_bytes(_lib.OP_DENSITY_FWD, N, n, flags, 0))
        one = jax.ShapeDtypeStruct((1,), f64)
        outs = _call("density_fwd", (jax.ShapeDtypeStruct((N, 2), f64), jax.ShapeDtypeStruct((N, 2, 3), f64), one, one, one,
                                     jax.ShapeDtypeStruct((ws,), jnp.uint8)),
                     packed, rdm1, jnp.zeros((1,), f64), opaque=pack_dims(N=N, n=n, flags=flags, nplanes=nplanes, ws_bytes=ws))
        return outs[0], outs[1]

    def fwd_rule(rdm1):
        return fwd_op(rdm1), None

    def bwd_rule(_, cot):
        rb, gb = cot
        ws = int(L.gdft_workspace_bytes(_lib.OP_DENSITY_BWD, N, n, flags, 0))
        z = jnp.zeros((1,), f64)
        dbar, _ = _call("density_bwd", (jax.ShapeDtypeStruct((2, n, n), f64), jax.ShapeDtypeStruct((ws,), jnp.uint8)),
                        packed, rb, gb, z, z, opaque=pack_dims(N=N, n=n, flags=flags, nplanes=nplanes, ws_bytes=ws))
        return (dbar,)

    fwd_op.defvjp(fwd_rule, bwd_rule)
    return fwd_op(rdm1)
