"""A minimal stand-in for the few JAX symbols graddft_b200/jax_ffi.py touches (TEST INFRASTRUCTURE): `jax.custom_vjp`,
`jax.ShapeDtypeStruct`, `jax.ffi.{ffi_call, register_ffi_target, pycapsule}` and a torch-backed `jax.numpy` subset.  With it
the JAX wrappers -- plans, operand order, residuals, forward and backward rules, the nesting that closes them under repeated
differentiation -- EXECUTE on the GPU without JAX: `ffi_call` hands the plan to the real XLA adapter through ctypes
(`jax_ffi.run_plan_torch`), exactly what XLA's thunk would do with the same buffers.  What it cannot stand in for is JAX's
tracing itself (jit / grad composition), which is JAX's own code, not this repository's.

`custom_vjp` objects record every call in CALLS, so a test can take the pullback of the call a wrapper made internally:
    out = jax_ffi.coulomb_j(eri, P);  f, args = shim.CALLS[-1];  primal, pull = shim.vjp(f, *args);  (Pbar,) = pull(Jbar)
"""
import types

import torch

CALLS = []


class custom_vjp:
    def __init__(self, fun):
        self.fun, self.fwd, self.bwd = fun, None, None

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd

    def __call__(self, *args):
        CALLS.append((self, args))
        return self.fun(*args)


def vjp(f: custom_vjp, *args):
    """(primal output, pullback) through the rules registered with defvjp, as jax.vjp would use them."""
    out, res = f.fwd(*args)
    return out, (lambda cot: f.bwd(res, cot))


class ShapeDtypeStruct:
    def __init__(self, shape, dtype):
        self.shape, self.dtype = tuple(shape), dtype


def _dtype_name(dt) -> str:
    return dt if isinstance(dt, str) else str(dt).replace("torch.", "")


def build(device):
    """(jax, jax.numpy) module objects whose arrays are torch tensors on `device`."""
    from graddft_b200 import jax_ffi

    jnp = types.ModuleType("jax.numpy")
    jnp.float64, jnp.uint8 = torch.float64, torch.uint8
    jnp.dtype = lambda d: getattr(torch, d) if isinstance(d, str) else d
    jnp.zeros = lambda shape, dtype=torch.float64: torch.zeros(shape, dtype=dtype, device=device)
    jnp.zeros_like = torch.zeros_like
    jnp.reshape = lambda x, shape: x.reshape(shape)
    jnp.swapaxes = lambda x, a, b: x.transpose(a, b)

    ffi = types.SimpleNamespace()
    ffi.registered = []
    ffi.register_ffi_target = lambda name, capsule, platform="CUDA", api_version=0: ffi.registered.append(name)
    ffi.pycapsule = lambda ptr: ptr

    def ffi_call(name, result_shape_dtypes, custom_call_api_version=2, legacy_backend_config=None, **_):
        assert name.startswith("gdft_") and custom_call_api_version in (1, 2) and isinstance(legacy_backend_config, bytes)

        def call(*operands):
            plan = jax_ffi.Plan(name[len("gdft_"):], len(operands), tuple((s.shape, _dtype_name(s.dtype)) for s in result_shape_dtypes),
                                legacy_backend_config)
            return tuple(jax_ffi.run_plan_torch(plan, list(operands)))

        return call

    ffi.ffi_call = ffi_call
    jax = types.ModuleType("jax")
    jax.custom_vjp, jax.ShapeDtypeStruct, jax.ffi, jax.numpy = custom_vjp, ShapeDtypeStruct, ffi, jnp
    return jax, jnp
