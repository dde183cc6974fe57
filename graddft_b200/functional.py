"""`Functional`, `NeuralFunctional`, `DM21` and the DM21 feature functions of grad_dft/functional.py.

Same names and argument meaning as the reference (grad_dft/functional.py:48-342 for `Functional`,
:345-498 `NeuralFunctional`, :504-758 the DM21 feature/combination functions, :761-822 `DM21`), with
torch tensors and `params` as a plain dict pytree of tensors (``{"params": {...}}`` accepted too).  The
grid-sized arithmetic (quadrature, closed-form features) is C-ABI kernel calls through `ops`; the
coefficient network is user-level host-framework code (float64 torch matmuls -> cuBLAS), as upstream where
it is flax code.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Sequence

import torch

from . import ops
from .molecule import Grid, Molecule, abs_clip

Array = torch.Tensor
F64 = torch.float64


def stop_gradient(x: Array) -> Array:
    return x.detach()


def _is_closed_form_row(c: Array) -> bool:
    return c.dim() == 2 and c.shape[0] == 1


@dataclass
class Functional:
    """grad_dft/functional.py:48-342.  `coefficients(self, cinputs)` and the feature callables keep the
    reference's calling conventions; `needs` is an optional hint (not upstream) naming the grid
    quantities the feature functions will ask the molecule for, so that they are produced by one fused
    kernel launch instead of one launch each."""

    coefficients: Callable
    energy_densities: Optional[Callable]
    coefficient_inputs: Optional[Callable] = None
    nograd_densities: Optional[Callable] = None
    densitygrads: Optional[Callable] = None
    combine_densities: Optional[Callable] = None
    nograd_coefficient_inputs: Optional[Callable] = None
    coefficient_input_grads: Optional[Callable] = None
    combine_inputs: Optional[Callable] = None
    is_xc: bool = True
    exchange_mask: Any = None
    needs: Sequence[str] = ()
    needs_omegas: Optional[Sequence[float]] = None

    # flax's Module.apply(params, *inputs): bind params, then __call__
    def apply(self, params, coefficient_inputs, **kwargs) -> Array:
        return self.coefficients(self, coefficient_inputs)

    def __call__(self, coefficient_inputs) -> Array:
        return self.coefficients(self, coefficient_inputs)

    def _prefetch(self, atoms) -> None:
        if self.needs and hasattr(atoms, "prefetch"):
            m = atoms._memo()
            from ._lib import GDFT_GRAD, GDFT_LAPL, GDFT_RHO, GDFT_TAU
            flag_of = {"rho": GDFT_RHO, "grad": GDFT_GRAD, "tau": GDFT_TAU, "lapl": GDFT_LAPL}
            if any(flag_of[nm] not in m for nm in self.needs):
                atoms.prefetch(*self.needs, omegas=self.needs_omegas)

    def compute_densities(self, atoms, clip_cte: float = 1e-30, *args, **kwargs) -> Array:
        """grad_dft/functional.py:160-185."""
        self._prefetch(atoms)
        if self.nograd_densities and self.energy_densities:
            densities = self.energy_densities(atoms, *args, **kwargs)
            nograd_densities = stop_gradient(self.nograd_densities(atoms, *args, **kwargs))
            densities = self.combine_densities(densities, nograd_densities)
        elif self.energy_densities:
            densities = self.energy_densities(atoms, *args, **kwargs)
        elif self.nograd_densities:
            densities = stop_gradient(self.nograd_densities(atoms, *args, **kwargs))
        return abs_clip(densities, clip_cte)

    def compute_coefficient_inputs(self, atoms, *args, **kwargs) -> Optional[Array]:
        """grad_dft/functional.py:187-217."""
        self._prefetch(atoms)
        if self.nograd_coefficient_inputs and self.coefficient_inputs:
            cinputs = self.coefficient_inputs(atoms, *args, **kwargs)
            nograd_cinputs = stop_gradient(self.nograd_coefficient_inputs(atoms, *args, **kwargs))
            cinputs = self.combine_inputs(cinputs, nograd_cinputs)
        elif self.coefficient_inputs:
            cinputs = self.coefficient_inputs(atoms, *args, **kwargs)
        elif self.nograd_coefficient_inputs:
            cinputs = stop_gradient(self.nograd_coefficient_inputs(atoms, *args, **kwargs))
        else:
            cinputs = None
        return cinputs

    def _features(self, atoms, tap: bool, clip_cte: float = 1e-30, *args, **kwargs) -> Dict[str, Any]:
        """compute_densities + compute_coefficient_inputs (functional.py:160-217) in one go, keeping the pieces.
        With `tap` the stop_gradient'ed pieces come back as fresh leaves ("tap_d", "tap_c"): differentiating E_xc
        w.r.t. them in the same backward pass that yields V_xc gives the cotangents arriving at the stop_gradient
        boundary, which is what the explicit exact-exchange routes (molecule.py:545-685) ask for again later."""
        self._prefetch(atoms)

        def stopped(t):
            t = stop_gradient(t)
            return t.requires_grad_(True) if tap else t

        ft: Dict[str, Any] = dict(grad_densities=None, tap_d=None, grad_cinputs=None, tap_c=None, cinputs=None)
        if self.energy_densities:
            ft["grad_densities"] = self.energy_densities(atoms, *args, **kwargs)
        if self.nograd_densities:
            ft["tap_d"] = stopped(self.nograd_densities(atoms, *args, **kwargs))
        if ft["grad_densities"] is not None and ft["tap_d"] is not None:
            raw = self.combine_densities(ft["grad_densities"], ft["tap_d"])
        else:
            raw = ft["grad_densities"] if ft["grad_densities"] is not None else ft["tap_d"]
        ft["densities_raw"] = raw
        ft["densities"] = abs_clip(raw, clip_cte)
        if self.coefficient_inputs:
            ft["grad_cinputs"] = self.coefficient_inputs(atoms, *args, **kwargs)
        if self.nograd_coefficient_inputs:
            ft["tap_c"] = stopped(self.nograd_coefficient_inputs(atoms, *args, **kwargs))
        if ft["grad_cinputs"] is not None and ft["tap_c"] is not None:
            ft["cinputs"] = self.combine_inputs(ft["grad_cinputs"], ft["tap_c"])
        else:
            ft["cinputs"] = ft["grad_cinputs"] if ft["grad_cinputs"] is not None else ft["tap_c"]
        return ft

    def coefficients_for(self, params, coefficient_inputs, densities: Array, **kwargs) -> Array:
        """The coefficient block c[r|1, f] xc_energy contracts with the densities (functional.py:246-250)."""
        coefficients = self.apply(params, coefficient_inputs, **kwargs)
        if coefficients.dim() == 1:
            coefficients = coefficients.unsqueeze(-1) if coefficients.shape[0] == densities.shape[0] else coefficients.unsqueeze(0)
        if coefficients.shape[-1] == 1 and densities.shape[1] != 1:
            coefficients = coefficients.expand(coefficients.shape[0], densities.shape[1])  # einsum "rf,rf->r" broadcast
        return coefficients.to(densities.dtype)

    def xc_energy(self, params, grid: Grid, coefficient_inputs, densities: Array, clip_cte: float = 1e-30, **kwargs) -> Array:
        """grad_dft/functional.py:219-253: e_r = sum_f c[r,f] d[r,f]; abs_clip; quadrature with clipped weights.
        One fused kernel (gdft_xc_integrate_fwd); its VJP is gdft_xc_integrate_bwd."""
        return ops.xc_integrate(self.coefficients_for(params, coefficient_inputs, densities, **kwargs), densities, grid.weights, clip_cte)

    def energy(self, params, atoms, *args, **kwargs) -> Array:
        """grad_dft/functional.py:255-288."""
        densities = self.compute_densities(atoms, *args, **kwargs)
        cinputs = self.compute_coefficient_inputs(atoms, *args)
        energy = self.xc_energy(params, atoms.grid, cinputs, densities, **kwargs)
        if self.is_xc:
            energy = energy + atoms.nonXC()
        return energy

    def energy_xc_only(self, params, atoms, *args, **kwargs) -> Array:
        """grad_dft/functional.py:290-314."""
        densities = self.compute_densities(atoms, *args, **kwargs)
        cinputs = self.compute_coefficient_inputs(atoms, *args)
        return self.xc_energy(params, atoms.grid, cinputs, densities, **kwargs)

    def _integrate(self, energy_density: Array, gridweights: Array, precision=None, clip_cte: float = 1e-30) -> Array:
        """grad_dft/functional.py:316-342 (kept for constraint-style callers that integrate their own density)."""
        ones = torch.ones((1, 1), dtype=energy_density.dtype, device=energy_density.device)
        return ops.xc_integrate(ones, energy_density.reshape(-1, 1), gridweights, clip_cte)


# ---------------------------------------------------------------------------------------------------------
# neural functionals
# ---------------------------------------------------------------------------------------------------------
def _unwrap(params):
    return params["params"] if isinstance(params, dict) and "params" in params and isinstance(params["params"], dict) else params


@dataclass
class NeuralFunctional(Functional):
    """grad_dft/functional.py:345-498.  `coefficients(self, cinputs)` may call `self.dense(i, x)`,
    `self.layer_norm(i, x)` and `self.head(x, ...)`, which read the bound parameter dict."""

    activation: Callable = torch.nn.functional.gelu
    param_dtype: torch.dtype = F64

    def apply(self, params, coefficient_inputs, **kwargs) -> Array:
        object.__setattr__(self, "_bound", _unwrap(params))
        object.__setattr__(self, "_dense_i", 0)
        object.__setattr__(self, "_ln_i", 0)
        try:
            return self.coefficients(self, coefficient_inputs)
        finally:
            object.__setattr__(self, "_bound", None)

    # flax names submodules Dense_0, Dense_1, ... / LayerNorm_0, ... in call order
    def dense(self, x: Array) -> Array:
        i = self._dense_i
        object.__setattr__(self, "_dense_i", i + 1)
        p = self._bound
        kernel, bias = p[f"Dense_{i}.kernel"], p[f"Dense_{i}.bias"]
        if ops.dense_supported(x, kernel):
            return ops.dense_layer(x, kernel, bias)  # FP64 tensor-core GEMM of the library (gdft_dense_fwd)
        return x @ kernel + bias  # second-order builds, CPU tensors of the tests' oracles, widths beyond the kernel

    def layer_norm(self, x: Array, eps: float = 1e-6) -> Array:
        i = self._ln_i
        object.__setattr__(self, "_ln_i", i + 1)
        p = self._bound
        mu = x.mean(dim=-1, keepdim=True)
        var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
        return (x - mu) * torch.rsqrt(var + eps) * p[f"LayerNorm_{i}.scale"] + p[f"LayerNorm_{i}.bias"]

    def residual_block(self, x: Array, eps: float = 1e-6) -> Array:
        """One loop body of default_nn (grad_dft/functional.py:809-819): activation(LayerNorm(dense(x) + x)).  With the ELU
        activation inside a first-order build this is the library GEMM followed by ONE fused kernel pass
        (gdft_ln_elu_fwd / _bwd) instead of ~10 elementwise kernels; otherwise the same composite as upstream."""
        if self.activation is torch.nn.functional.elu and ops.residual_trunk_supported(x, x.shape[-1]):
            return self.residual_trunk(x, 1, eps)
        if self.activation is torch.nn.functional.elu and ops.residual_layernorm_elu_supported(x):
            i = self._ln_i
            object.__setattr__(self, "_ln_i", i + 1)
            p = self._bound
            j = self._dense_i
            object.__setattr__(self, "_dense_i", j + 1)
            # the Dense bias is added inside the fused pass (and its cotangent summed there): y is the bare GEMM
            return ops.residual_layernorm_elu(x @ p[f"Dense_{j}.kernel"], x, p[f"LayerNorm_{i}.scale"], p[f"LayerNorm_{i}.bias"], eps,
                                              ybias=p[f"Dense_{j}.bias"])
        return self.activation(self.layer_norm(self.dense(x) + x, eps))

    def residual_trunk(self, x: Array, nblocks: int, eps: float = 1e-6) -> Array:
        """`nblocks` consecutive residual blocks (the whole loop of default_nn, functional.py:809-819).  Inside a first-order
        build with the ELU activation and a width the GEMM kernels take, the run is one differentiable unit of library
        kernels (ops.residual_trunk: one fused GEMM per block forward, one per block reverse); otherwise block by block."""
        if not (self.activation is torch.nn.functional.elu and ops.residual_trunk_supported(x, x.shape[-1])):
            for _ in range(nblocks):
                x = self.residual_block(x, eps)
            return x
        p = self._bound
        i, j = self._ln_i, self._dense_i
        blocks = []
        for b in range(nblocks):
            kernel = p[f"Dense_{j + b}.kernel"]
            if tuple(kernel.shape) != (x.shape[-1], x.shape[-1]):
                raise ValueError("residual blocks need square Dense kernels of the input width")
            blocks.append((kernel, p[f"Dense_{j + b}.bias"], p[f"LayerNorm_{i + b}.scale"], p[f"LayerNorm_{i + b}.bias"]))
        object.__setattr__(self, "_ln_i", i + nblocks)
        object.__setattr__(self, "_dense_i", j + nblocks)
        return ops.residual_trunk(x, blocks, eps)

    def head(self, x: Array, local_features: int, sigmoid_scale_factor: float) -> Array:
        """grad_dft/functional.py:407-419: dense -> sigmoid(x/s) * s."""
        x = self.dense(x)
        return sigmoid_scale_factor * torch.sigmoid(x / sigmoid_scale_factor)


def canonicalize_inputs(x):
    """grad_dft/functional.py:931-947."""
    if isinstance(x, (tuple, list)):
        x = torch.cat(list(x), dim=1)
    if x.dim() == 1:
        x = x.unsqueeze(1)
    return x


# ---- DM21 features (grad_dft/functional.py:504-758) -------------------------------------------------
def dm21_coefficient_inputs(molecule: Molecule, clip_cte: float = 1e-30, *_, **__) -> Array:
    """grad_dft/functional.py:504-531: [rho_a, rho_b, |g_a+g_b|^2, |g_a|^2, |g_b|^2, tau_a, tau_b]."""
    return ops.pointwise("DM21_INPUTS", molecule.density(), molecule.grad_density(), molecule.kinetic_density(), None, clip_cte)


def dm21_densities(molecule: Molecule, functional_type: Optional[str] = "LDA", clip_cte: float = 1e-30, *_, **__) -> Array:
    """grad_dft/functional.py:534-626."""
    kind = {"LDA": "DM21_LDA", "DM21": "DM21_LDA", "GGA": "DM21_GGA", "MGGA": "DM21_MGGA"}[functional_type]
    rho = molecule.density()
    grho = molecule.grad_density() if kind != "DM21_LDA" else None
    tau = molecule.kinetic_density() if kind == "DM21_MGGA" else None
    return ops.pointwise(kind, rho, grho, tau, None, clip_cte)


def densities(molecule: Molecule, functional_type: Optional[str] = "LDA", clip_cte: float = 1e-30, *_, **__) -> Array:
    """grad_dft/functional.py:1048-1202: the u/w enhancement-factor feature library the article's neural functionals
    train on.  Per-spin exchange columns rho^{4/3} u^i w^j followed by the same number of correlation columns, which
    are identically zero upstream (jnp.round(e_PW92, -30) == 0) and are reproduced as zeros with
    zero gradient.  As upstream, only the three string options work (functional.py:1087-1098)."""
    if not isinstance(functional_type, str) or functional_type not in ("LDA", "DM21", "GGA", "MGGA"):
        raise ValueError(f"Functional type {functional_type} not recognized, must be one of LDA, GGA, MGGA.")
    kind = {"LDA": "FEAT_LDA", "DM21": "FEAT_LDA", "GGA": "FEAT_GGA", "MGGA": "FEAT_MGGA"}[functional_type]
    rho = molecule.density()
    grho = molecule.grad_density() if kind != "FEAT_LDA" else None
    tau = molecule.kinetic_density() if kind == "FEAT_MGGA" else None
    return ops.pointwise(kind, rho, grho, tau, None, clip_cte)


def dm21_combine_cinputs(cinputs: Array, ehf: Array) -> Array:
    """grad_dft/functional.py:628-649: HF features appended by spin, [w0 a, w1 a, w0 b, w1 b]."""
    return torch.cat([cinputs, ehf[:, 0].T, ehf[:, 1].T], dim=1)


def dm21_combine_densities(densities: Array, ehf: Array) -> Array:
    """grad_dft/functional.py:651-675: one spin-summed HF column per omega."""
    return torch.cat([densities] + [ehf[i].sum(dim=0, keepdim=True).T for i in range(ehf.shape[0])], dim=1)


def dm21_hfgrads_densities(functional, params, atoms, ehf, coefficient_inputs, densities_wout_hf, omegas=(0.0, 0.4)) -> Array:
    """grad_dft/functional.py:677-717."""
    vxc_hf = atoms.HF_density_grad_2_Fock(functional, params, omegas, ehf, coefficient_inputs, densities_wout_hf)
    return vxc_hf.sum(dim=0)


def dm21_hfgrads_cinputs(functional, params, atoms, ehf, cinputs_wout_hf, densities, omegas=(0.0, 0.4)) -> Array:
    """grad_dft/functional.py:719-758."""
    vxc_hf = atoms.HF_coefficient_input_grad_2_Fock(functional, params, omegas, ehf, cinputs_wout_hf, densities)
    return vxc_hf.sum(dim=0)


_DM21_OMEGAS = (0.0, 0.4)


def standard_hf_routes(functional) -> Optional[Sequence[float]]:
    """The omegas of a functional whose explicit exact-exchange routes are the built-in ones -- V = sum_w hf_fock(dE_xc/d e_HF)[w]
    through the densities and/or through the coefficient inputs (grad_dft/functional.py:677-758, popular_functionals.py:357-372)
    -- else None.  For those the predictor may add the two cotangents BEFORE the GEMM (the map g -> V is linear) and take the
    sum over omega inside it: one two-unit GEMM instead of 2 x len(omegas); the only difference to upstream's term-by-term
    `fock += V + V^T; abs_clip` is the clip of an intermediate sum at 1e-30."""
    from . import popular_functionals as pf

    if isinstance(functional, DM21):
        fields = DM21.__dataclass_fields__
        same = (functional.densitygrads is fields["densitygrads"].default and functional.coefficient_input_grads is fields["coefficient_input_grads"].default
                and functional.nograd_densities is fields["nograd_densities"].default
                and functional.nograd_coefficient_inputs is fields["nograd_coefficient_inputs"].default)
        return _DM21_OMEGAS if same else None
    if functional is pf.B3LYP:
        return (0.0,)
    return None


def _dm21_default_nn(instance, rhoinputs, *_, **__):
    """grad_dft/functional.py:793-822: log|x|+eps -> dense -> tanh -> 6 x (dense + res -> LayerNorm -> act) -> head."""
    x = canonicalize_inputs(rhoinputs)
    x = torch.log(torch.abs(x) + instance.squash_offset)
    x = torch.tanh(instance.dense(x))
    x = instance.residual_trunk(x, len(instance.layer_widths))
    return instance.head(x, instance.local_features, instance.sigmoid_scale_factor)


@dataclass
class DM21(NeuralFunctional):
    """grad_dft/functional.py:761-928: architecture, feature wiring and `generate_DM21_weights` (the published weights
    read from the TF checkpoint bundle without TensorFlow, or seeded initial values)."""

    coefficients: Callable = _dm21_default_nn
    energy_densities: Callable = dm21_densities
    nograd_densities: Callable = lambda atoms, *_, **__: atoms.HF_energy_density(_DM21_OMEGAS)
    densitygrads: Callable = lambda self, params, atoms, nograd_densities, cinputs, grad_densities, *_, **__: dm21_hfgrads_densities(
        self, params, atoms, nograd_densities, cinputs, grad_densities, _DM21_OMEGAS)
    combine_densities: Callable = dm21_combine_densities
    coefficient_inputs: Callable = dm21_coefficient_inputs
    nograd_coefficient_inputs: Callable = lambda atoms, *_, **__: atoms.HF_energy_density(_DM21_OMEGAS)
    coefficient_input_grads: Callable = lambda self, params, atoms, nograd_cinputs, grad_cinputs, densities, *_, **__: dm21_hfgrads_cinputs(
        self, params, atoms, nograd_cinputs, grad_cinputs, densities, _DM21_OMEGAS)
    combine_inputs: Callable = dm21_combine_cinputs
    activation: Callable = torch.nn.functional.elu
    squash_offset: float = 1e-4
    layer_widths: Sequence[int] = (256, 256, 256, 256, 256, 256)
    local_features: int = 3
    sigmoid_scale_factor: float = 2.0
    needs: Sequence[str] = ("rho", "grad", "tau")
    needs_omegas: Optional[Sequence[float]] = _DM21_OMEGAS

    def default_nn(self, rhoinputs, *a, **k):
        return _dm21_default_nn(self, rhoinputs)

    def generate_DM21_weights(self, folder: Optional[str] = None, num_layers_with_dm_parameters: int = 7, n_input_features: int = 11,
                              seed: int = 1984, device=None) -> Dict[str, Array]:
        """grad_dft/functional.py:824-928.  With `folder` (e.g. "models/DM21_model" of a Grad DFT checkout, or one of the
        DeepMind `checkpoints/DM21*` folders) the weights of the published DM21 network are read straight from the TF
        checkpoint bundle (`tf_bundle.load_variables`; upstream goes through `tf.saved_model.load`, which needs TensorFlow)
        and merged into freshly initialised parameters by upstream's rule: a layer takes the checkpoint's arrays when its
        index is <= `num_layers_with_dm_parameters` and every shape agrees, else it keeps its initial values (plus the
        identity for square Dense kernels, functional.py:913-921).  Variable -> parameter naming as in upstream's
        `vars_to_params` (functional.py:861-893): SquashUnprocessedData -> Dense_0, ResidualBlock[_k] -> Dense_{k+1} and
        LayerNorm_k, OutputLayer -> Dense_7.  The checkpoint is float32; values are widened to float64 exactly.
        Without `folder`: seeded initial values only (He-normal kernels, zero biases, LayerNorm scale 1 / bias 0)."""
        g = torch.Generator().manual_seed(seed)
        widths = list(self.layer_widths)
        p: Dict[str, Array] = {}

        def he(i, o):
            return torch.randn(i, o, generator=g, dtype=F64) * math.sqrt(2.0 / i)

        p["Dense_0.kernel"], p["Dense_0.bias"] = he(n_input_features, widths[0]), torch.zeros(widths[0], dtype=F64)
        for k, wdt in enumerate(widths):
            p[f"Dense_{k + 1}.kernel"] = he(wdt, wdt)
            p[f"Dense_{k + 1}.bias"] = torch.zeros(wdt, dtype=F64)
            p[f"LayerNorm_{k}.scale"] = torch.ones(wdt, dtype=F64)
            p[f"LayerNorm_{k}.bias"] = torch.zeros(wdt, dtype=F64)
        p[f"Dense_{len(widths) + 1}.kernel"] = he(widths[-1], self.local_features)
        p[f"Dense_{len(widths) + 1}.bias"] = torch.zeros(self.local_features, dtype=F64)

        dm = dm21_checkpoint_params(folder) if folder is not None else {}
        modules = sorted({k.split(".")[0] for k in p})
        for mod in modules:
            leaves = [k for k in p if k.split(".")[0] == mod]
            same = all(k in dm and tuple(dm[k].shape) == tuple(p[k].shape) for k in leaves)
            if int(mod.split("_")[1]) > num_layers_with_dm_parameters or not same:
                kern = p.get(f"{mod}.kernel")
                if mod.startswith("Dense") and kern.shape[0] == kern.shape[1]:
                    p[f"{mod}.kernel"] = kern + torch.eye(kern.shape[0], dtype=F64)
            else:
                for k in leaves:
                    p[k] = dm[k]
        if device is not None:
            p = {k: v.to(device) for k, v in p.items()}
        return p


def dm21_checkpoint_params(folder: str) -> Dict[str, Array]:
    """The variables of a DM21 TF checkpoint under the flax-style names of grad_dft/functional.py:861-893."""
    import re

    from . import tf_bundle

    if not os.path.isabs(folder) and not os.path.isdir(folder):
        raise FileNotFoundError(f"DM21 checkpoint folder {folder!r} not found")
    out: Dict[str, Array] = {}
    for name, arr in tf_bundle.load_variables(folder).items():
        if "ResidualBlock_" in name:
            number = int(re.findall("ResidualBlock_[0-9]", name)[0][-1]) + 1
        elif "ResidualBlock/" in name:
            number = 1
        elif "Squash" in name:
            number = 0
        elif "Output" in name:
            number = 7
        else:
            raise ValueError(f"Unknown variable name {name!r}.")
        t = torch.from_numpy(arr.astype("float64"))
        if "/linear/" in name:
            if name.endswith("/w"):
                out[f"Dense_{number}.kernel"] = t
            elif name.endswith("/b"):
                out[f"Dense_{number}.bias"] = t
        elif "/layer_norm/" in name:
            if name.endswith("gamma"):
                out[f"LayerNorm_{number - 1}.scale"] = t
            elif name.endswith("beta"):
                out[f"LayerNorm_{number - 1}.bias"] = t
    return out
