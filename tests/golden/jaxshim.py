"""A torch-backed stand-in for the handful of jax / flax / jaxtyping symbols that the reference's hot-path
modules use, so that the reference's OWN, UNMODIFIED source files (grad_dft/molecule.py, functional.py,
popular_functionals.py, train.py under /root/reference) can be imported and executed in a container that
has no jax.  TEST INFRASTRUCTURE, used only by tests/golden/make_golden.py to generate the committed golden
vectors (the reference cannot travel to the GPU box; the vectors can).

What this is and is not: the arithmetic EXPRESSIONS that run are the reference's (its einsum strings, its
clip/where guards, its log2/exp2-domain algebra, its value_and_grad / grad call structure, its predictor
assembly order); the array BACKEND that evaluates them is torch float64 on CPU instead of XLA.  jnp.where /
jnp.clip map to torch.where / torch.clamp, whose (sub)gradient conventions coincide with JAX's away from exact
ties.  `vmap_chunked` (grad_dft/utils/chunk.py, a memory-limiting device built on jax internals) is replaced by
a plain vmap; `jax.jit` is the identity.
"""
from __future__ import annotations

import dataclasses
import enum
import importlib
import importlib.util
import math
import sys
import types
from functools import partial
from pathlib import Path

import torch

F64 = torch.float64
REF = Path("/root/reference")


# ---------------------------------------------------------------------------------------------------------
# array type: torch.Tensor with the numpy-flavoured methods the reference calls
# ---------------------------------------------------------------------------------------------------------
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        out = self.arr.clone()
        idx = self.idx
        # JAX scatter semantics: out-of-bounds updates are dropped
        parts = idx if isinstance(idx, tuple) else (idx,)
        if len(parts) == 2 and all(isinstance(p, torch.Tensor) for p in parts):  # diag_indices_from
            out[parts[0], parts[1]] = value
            return out.as_subclass(JArray)
        for d, p in enumerate(parts):
            if isinstance(p, int) and not (-out.shape[d] <= p < out.shape[d]):
                return out.as_subclass(JArray)
        out[idx] = value
        return out.as_subclass(JArray)


class JArray(torch.Tensor):
    def transpose(self, *axes):
        if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
            axes = tuple(axes[0])
        if len(axes) == 0:
            axes = tuple(reversed(range(self.dim())))
        return self.permute(*axes)

    @property
    def at(self):
        return _At(self)

    def __getitem__(self, idx):
        # JAX gather semantics: out-of-bounds integer indices are clamped
        parts = idx if isinstance(idx, tuple) else (idx,)
        if any(isinstance(p, int) for p in parts):
            fixed, d = [], 0
            for p in parts:
                if p is None:
                    fixed.append(p)
                    continue
                if p is Ellipsis:
                    fixed.append(p)
                    d = self.dim() - (len([q for q in parts[parts.index(p) + 1:] if q is not None]))
                    continue
                if isinstance(p, int) and self.dim() > d:
                    n = self.shape[d]
                    p = min(max(p, -n), n - 1)
                fixed.append(p)
                d += 1
            idx = tuple(fixed) if isinstance(idx, tuple) else fixed[0]
        return super().__getitem__(idx)

    def astype(self, dt):
        return self.to(dt)

    def __contains__(self, item):
        return bool((self == item).any())


def _j(x):
    if isinstance(x, JArray):
        return x
    if isinstance(x, torch.Tensor):
        return x.as_subclass(JArray)
    return torch.as_tensor(x, dtype=F64 if isinstance(x, float) or (isinstance(x, (list, tuple)) and _has_float(x)) else None).as_subclass(JArray)


def _has_float(x):
    if isinstance(x, (list, tuple)):
        return any(_has_float(v) for v in x)
    return isinstance(x, float)


# ---------------------------------------------------------------------------------------------------------
# jax.numpy
# ---------------------------------------------------------------------------------------------------------
jnp = types.ModuleType("jax.numpy")
jnp.pi = math.pi


class _DT:
    """jnp.int64 & co: usable as a dtype and callable as a scalar constructor."""

    def __init__(self, dt, py):
        self.dt, self.py = dt, py

    def __call__(self, x):
        return self.py(x)


jnp.float64, jnp.float32, jnp.int64, jnp.int32 = _DT(torch.float64, float), _DT(torch.float32, float), _DT(torch.int64, int), _DT(torch.int32, int)
jnp.ndarray = torch.Tensor
jnp.newaxis = None


def _einsum(spec, *ops, precision=None, **kw):
    return torch.einsum(spec.replace(" ", ""), *[_j(o) for o in ops])


def _array(x, dtype=None):
    dtype = getattr(dtype, "dt", dtype)
    if isinstance(x, torch.Tensor):
        return _j(x if dtype is None else x.to(dtype))
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return _j(torch.stack(list(x)))
    t = torch.as_tensor(x, dtype=F64) if _has_float(x) else torch.as_tensor(x)
    if t.is_floating_point():
        t = t.to(F64)
    if dtype is not None:
        t = t.to(dtype)
    return _j(t)


def _where(c, a, b):
    c = _j(c)
    ref = a if isinstance(a, torch.Tensor) else (b if isinstance(b, torch.Tensor) else None)
    dt = ref.dtype if ref is not None else F64
    if not isinstance(a, torch.Tensor):
        a = torch.as_tensor(a, dtype=dt)
    if not isinstance(b, torch.Tensor):
        b = torch.as_tensor(b, dtype=dt)
    return _j(torch.where(c, a, b))


def _clip(x, a_min=None, a_max=None):
    return _j(torch.clamp(_j(x), min=a_min, max=a_max))


def _sum(x, axis=None, keepdims=False):
    x = _j(x)
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def _round(x, decimals=0):
    # numpy semantics: round(x, d) = rint(x * 10^d) / 10^d ; zero gradient
    s = 10.0 ** decimals
    return _j(torch.round(_j(x).detach() * s) / s)


jnp.einsum = _einsum
jnp.array = _array
jnp.asarray = _array
jnp.where = _where
jnp.clip = _clip
jnp.sum = _sum
jnp.round = _round
jnp.abs = lambda x: _j(torch.abs(_j(x)))
jnp.log2 = lambda x: _j(torch.log2(_j(x)))
jnp.log = lambda x: _j(torch.log(_j(x)))
jnp.exp = lambda x: _j(torch.exp(_j(x)))
jnp.sqrt = lambda x: _j(torch.sqrt(_j(x)))
jnp.arctan = lambda x: _j(torch.atan(_j(x)))
jnp.arcsinh = lambda x: _j(torch.asinh(_j(x)))
jnp.tanh = lambda x: _j(torch.tanh(_j(x)))
jnp.sign = lambda x: _j(torch.sign(_j(x)))
jnp.maximum = lambda a, b: _j(torch.maximum(_j(a), torch.as_tensor(b, dtype=_j(a).dtype)))
jnp.minimum = lambda a, b: _j(torch.minimum(_j(a), torch.as_tensor(b, dtype=_j(a).dtype)))
jnp.stack = lambda xs, axis=0: _j(torch.stack([_j(x) for x in xs], dim=axis))
jnp.concatenate = lambda xs, axis=0: _j(torch.cat([_j(x) for x in xs], dim=axis))
jnp.expand_dims = lambda x, axis: _j(_j(x).unsqueeze(axis))
jnp.squeeze = lambda x, axis=None: _j(_j(x).squeeze() if axis is None else _j(x).squeeze(axis))
jnp.zeros_like = lambda x, dtype=None: _j(torch.zeros_like(_j(x), dtype=getattr(dtype, "dt", dtype)))
jnp.ones_like = lambda x, dtype=None: _j(torch.ones_like(_j(x), dtype=getattr(dtype, "dt", dtype)))
jnp.zeros = lambda shape, dtype=F64: _j(torch.zeros(shape, dtype=getattr(dtype, "dt", dtype)))
jnp.ones = lambda shape, dtype=F64: _j(torch.ones(shape, dtype=getattr(dtype, "dt", dtype)))
jnp.eye = lambda n, dtype=F64: _j(torch.eye(n, dtype=getattr(dtype, "dt", dtype)))
jnp.arange = lambda *a, **k: _j(torch.arange(*a, **k))
jnp.tensordot = lambda a, b, axes: _j(torch.tensordot(_j(a), _j(b), dims=([axes[0]] if isinstance(axes[0], int) else list(axes[0]),
                                                                       [axes[1]] if isinstance(axes[1], int) else list(axes[1]))))
jnp.isnan = lambda x: torch.isnan(_j(x))
jnp.isinf = lambda x: torch.isinf(_j(x))
jnp.less = lambda a, b: a < b
jnp.greater = lambda a, b: _j(a) > b
jnp.argsort = lambda x: torch.argsort(_j(x), stable=True)
jnp.power = lambda a, b: _j(torch.pow(_j(a), b))
jnp.mean = lambda x, axis=None: _j(_j(x).mean() if axis is None else _j(x).mean(dim=axis))
jnp.empty = lambda shape, dtype=F64: _j(torch.empty(shape, dtype=getattr(dtype, "dt", dtype)))
jnp.dot = lambda a, b: _j(_j(a) @ _j(b))
jnp.logical_and = lambda a, b: torch.logical_and(torch.as_tensor(a), torch.as_tensor(b))
jnp.reshape = lambda x, shape: _j(_j(x).reshape(shape))
jnp.identity = lambda n, dtype=F64: _j(torch.eye(n, dtype=getattr(dtype, "dt", dtype)))
jnp.inf = float("inf")
jnp.moveaxis = lambda x, a, b: _j(torch.movedim(_j(x), a, b))
jnp.vectorize = lambda f, signature=None: f
jnp.nan_to_num = lambda x: _j(torch.nan_to_num(_j(x)))
jnp.divide = lambda a, b: _j(torch.as_tensor(a, dtype=F64) / b)
jnp.multiply = lambda a, b: _j(a * b)
jnp.diag_indices_from = lambda x: (torch.arange(x.shape[0]), torch.arange(x.shape[0]))
jnp.less_equal = lambda a, b: a <= b
jnp.isclose = lambda a, b, **k: torch.isclose(torch.as_tensor(a, dtype=F64), torch.as_tensor(b, dtype=F64), **k)
jnp.allclose = lambda a, b, **k: bool(torch.allclose(torch.as_tensor(a, dtype=F64), torch.as_tensor(b, dtype=F64), **k))
jnp.hstack = lambda xs: _j(torch.hstack([_j(x) for x in xs]))
jnp.any = lambda x: bool(torch.as_tensor(x).any())
jnp.diag = lambda x: _j(torch.diag(_j(x)))
jnp.linalg = types.SimpleNamespace(
    norm=lambda x, ord=None: _j(torch.linalg.norm(_j(x))), eigh=lambda x: tuple(_j(t) for t in torch.linalg.eigh(_j(x))),
    inv=lambda x: _j(torch.linalg.inv(_j(x))), cholesky=lambda x: _j(torch.linalg.cholesky(_j(x))),
)


# ---------------------------------------------------------------------------------------------------------
# jax core transforms
# ---------------------------------------------------------------------------------------------------------
def jit(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def _leafify(x):
    if isinstance(x, torch.Tensor):
        base = x if x.requires_grad else x.detach()
        return _j(base.clone().requires_grad_(True)) if not x.requires_grad else x
    return _j(torch.tensor(float(x), dtype=F64, requires_grad=True))


def value_and_grad(f, argnums=0, has_aux=False):
    def wrapped(*args, **kw):
        args = list(args)
        single = isinstance(argnums, int)
        idx = (argnums,) if single else tuple(argnums)
        leaves = []
        for i in idx:
            args[i] = _leafify(args[i])
            leaves.append(args[i])
        with torch.enable_grad():
            out = f(*args, **kw)
            val, aux = (out if has_aux else (out, None))
            grads = torch.autograd.grad(val, leaves, create_graph=True, allow_unused=True)
        grads = tuple(_j(g) if g is not None else jnp.zeros_like(l) for g, l in zip(grads, leaves))
        g = grads[0] if single else grads
        return ((val, aux), g) if has_aux else (val, g)

    return wrapped


def grad(f, argnums=0, has_aux=False):
    vg = value_and_grad(f, argnums, has_aux)

    def wrapped(*args, **kw):
        (v, g) = vg(*args, **kw)
        return (g, v[1]) if has_aux else g

    return wrapped


def vmap(f, in_axes=0, out_axes=0):
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(a.shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = []
        for i in range(n):
            sl = [a if ax is None else _j(a).select(ax, i) for a, ax in zip(args, axes)]
            outs.append(f(*sl))
        if isinstance(outs[0], tuple):
            return tuple(_j(torch.stack([o[k] for o in outs], dim=out_axes)) for k in range(len(outs[0])))
        return _j(torch.stack(outs, dim=out_axes))

    return wrapped


def vmap_chunked(f, in_axes=0, *, chunk_size=None):
    """stand-in for grad_dft/utils/chunk.py: a vectorised map; chunking only bounds memory."""
    return vmap(f, in_axes=in_axes, out_axes=0)


class Precision(enum.Enum):
    DEFAULT = 0
    HIGH = 1
    HIGHEST = 2


def fori_loop(lower, upper, body_fun, init_val):
    val = init_val
    for i in range(int(lower), int(upper)):
        val = body_fun(i, val)
    return val


def cond(pred, t, f, *operands, **kw):
    if "operand" in kw:
        operands = (kw["operand"],)
    return t(*operands) if bool(pred) else f(*operands)


class custom_vjp:
    """forward values only: the golden vectors never differentiate through safe_eigh"""

    def __init__(self, f):
        self.f = f

    def __call__(self, *a, **k):
        return self.f(*a, **k)

    def defvjp(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd


# ---------------------------------------------------------------------------------------------------------
# flax stand-ins
# ---------------------------------------------------------------------------------------------------------
def struct_dataclass(cls):
    cls = dataclasses.dataclass(cls)
    cls.replace = lambda self, **kw: dataclasses.replace(self, **kw)
    return cls


_CTX = []


class Module:
    """flax.linen.Module stand-in: apply(params, *args) binds params and calls the module; submodules created
    inside the call are auto-named Class_i in creation order, as flax does."""

    def apply(self, params, *args, **kwargs):
        p = params["params"] if isinstance(params, dict) and "params" in params else params
        ctx = {"params": p, "count": {}}
        _CTX.append(ctx)
        try:
            if hasattr(self, "setup"):
                self.setup()
            return self(*args, **kwargs)
        finally:
            _CTX.pop()

    def sow(self, *a, **k):
        return None


def _next_params(kind):
    ctx = _CTX[-1]
    i = ctx["count"].get(kind, 0)
    ctx["count"][kind] = i + 1
    return ctx["params"][f"{kind}_{i}"]


class Dense:
    def __init__(self, features, **kw):
        self.features = features

    def __call__(self, x):
        p = _next_params("Dense")
        return _j(_j(x) @ p["kernel"] + p["bias"])


class LayerNorm:
    def __init__(self, epsilon=1e-6, **kw):
        self.eps = epsilon

    def __call__(self, x):
        p = _next_params("LayerNorm")
        x = _j(x)
        mu = x.mean(dim=-1, keepdim=True)
        var = (x * x).mean(dim=-1, keepdim=True) - mu * mu  # flax: fast variance E[x^2] - E[x]^2
        return _j((x - mu) * torch.rsqrt(var + self.eps) * p["scale"] + p["bias"])


class _Subscriptable:
    def __class_getitem__(cls, item):
        return cls


def _identity_decorator(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def install():
    """Register the stand-in modules in sys.modules and import the reference's hot-path modules."""
    if "grad_dft" in sys.modules and getattr(sys.modules["grad_dft"], "_shimmed", False):
        return sys.modules["grad_dft"]

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    lax = mod("jax.lax", Precision=Precision, stop_gradient=lambda x: _j(x.detach()) if isinstance(x, torch.Tensor) else x,
              fori_loop=fori_loop, cond=cond, map=lambda f, xs: _j(torch.stack([f(x) for x in xs])))
    nn_ = mod("jax.nn", sigmoid=lambda x: _j(torch.sigmoid(_j(x))), gelu=lambda x: _j(torch.nn.functional.gelu(_j(x), approximate="tanh")),
              elu=lambda x: _j(torch.nn.functional.elu(_j(x))))
    init = mod("jax.nn.initializers", zeros=None, he_normal=lambda *a, **k: None)
    nn_.initializers = init
    rnd = mod("jax.random", normal=None, PRNGKey=lambda s: s, split=lambda k, n=2: [k] * n)
    prof = mod("jax.profiler", annotate_function=lambda f=None, name=None, **k: f if f is not None else (lambda g: g))
    jsp_special = mod("jax.scipy.special", erfc=lambda x: _j(torch.erfc(_j(x))))
    jsp_opt = mod("jax.scipy.optimize", minimize=None)
    jsp = mod("jax.scipy", special=jsp_special, optimize=jsp_opt)
    tree_util = mod("jax.tree_util", tree_map=lambda f, t: {k: f(v) for k, v in t.items()} if isinstance(t, dict) else f(t),
                    tree_leaves=lambda t: list(t.values()) if isinstance(t, dict) else [t], tree_flatten=None)
    sys.modules["jax.numpy"] = jnp
    config = types.SimpleNamespace(x64_enabled=True, update=lambda *a, **k: None)
    jax = mod("jax", numpy=jnp, lax=lax, nn=nn_, random=rnd, profiler=prof, scipy=jsp, tree_util=tree_util, jit=jit, vmap=vmap,
              grad=grad, value_and_grad=value_and_grad, config=config, Array=torch.Tensor,
              custom_vjp=custom_vjp, debug=types.SimpleNamespace(print=lambda *a, **k: None))
    jt = mod("jaxtyping", jaxtyped=_identity_decorator)
    for nm in ("Array", "PyTree", "Scalar", "Float", "Int", "Complex", "PRNGKeyArray", "Bool"):
        setattr(jt, nm, type(nm, (_Subscriptable,), {}))
    mod("typeguard", typechecked=_identity_decorator)
    linen = mod("flax.linen", Module=Module, compact=lambda f: f, Dense=Dense, LayerNorm=LayerNorm)
    struct = mod("flax.struct", dataclass=struct_dataclass)
    core = mod("flax.core", freeze=lambda x: x, unfreeze=lambda x: x)
    ts = mod("flax.training.train_state", TrainState=object)
    training = mod("flax.training", train_state=ts, checkpoints=None)
    mod("flax", linen=linen, struct=struct, core=core, training=training)
    mod("optax", GradientTransformation=object, OptState=object, apply_updates=None)
    mod("chex")
    oc = mod("orbax.checkpoint", Checkpointer=object, PyTreeCheckpointer=object)
    mod("orbax", checkpoint=oc)

    # ---- the reference package, file by file (its __init__ pulls in pyscf/h5py-dependent modules) ----
    pkg = types.ModuleType("grad_dft")
    pkg.__path__ = [str(REF / "grad_dft")]
    pkg._shimmed = True
    sys.modules["grad_dft"] = pkg
    utypes = mod("grad_dft.utils.types", DType=object, default_dtype=lambda: F64, Array=torch.Tensor, PyTree=object, Scalar=object,
                 Hartree2kcalmol=627.50947)
    utils = mod("grad_dft.utils", vmap_chunked=vmap_chunked, types=utypes)
    utils.__path__ = []
    pkg.utils = utils

    def load(name):
        spec = importlib.util.spec_from_file_location(f"grad_dft.{name}", REF / "grad_dft" / f"{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"grad_dft.{name}"] = m
        spec.loader.exec_module(m)
        setattr(pkg, name, m)
        return m

    molecule = load("molecule")
    for nm in ("abs_clip", "Grid", "Molecule", "coulomb_energy", "coulomb_potential", "density", "grad_density", "lapl_density",
               "kinetic_density", "HF_energy_density", "nonXC", "make_rdm1", "get_occ", "orbital_grad", "one_body_energy"):
        setattr(pkg, nm, getattr(molecule, nm))
    pkg.Solid = type("Solid", (), {})
    functional = load("functional")
    for nm in ("Functional", "NeuralFunctional", "DM21", "DispersionFunctional", "correlation_polarization_correction",
               "exchange_polarization_correction", "dm21_coefficient_inputs", "dm21_densities", "densities",
               "dm21_combine_cinputs", "dm21_combine_densities"):
        setattr(pkg, nm, getattr(functional, nm))
    pop = load("popular_functionals")
    for nm in ("LSDA", "B88", "VWN", "LYP", "B3LYP", "PW92"):
        setattr(pkg, nm, getattr(pop, nm))
    train = load("train")
    pkg.energy_predictor = train.energy_predictor
    # SCF drivers: the eigen-solver and evaluate.py are the reference's; PySCF-only helpers are absent stubs
    spec = importlib.util.spec_from_file_location("grad_dft.utils.eigenproblem", REF / "grad_dft" / "utils" / "eigenproblem.py")
    eig = importlib.util.module_from_spec(spec)
    sys.modules["grad_dft.utils.eigenproblem"] = eig
    spec.loader.exec_module(eig)
    utils.safe_fock_solver = eig.safe_fock_solver
    utils.Optimizer = object
    iface = mod("grad_dft.interface", pyscf=mod("grad_dft.interface.pyscf", generate_chi_tensor=None, mol_from_Molecule=None, process_mol=None))
    iface.__path__ = []
    evaluate = load("evaluate")
    pkg.diff_scf_loop, pkg.diff_simple_scf_loop = evaluate.diff_scf_loop, evaluate.diff_simple_scf_loop
    pkg.J = _j
    return pkg
