"""LSDA, B88, VWN, LYP, B3LYP, PW92 -- grad_dft/popular_functionals.py, on the fused per-point kernels.

The closed-form energy densities (popular_functionals.py:29-269) are evaluated by `gdft_pointwise_fwd`
(their VJPs by `gdft_pointwise_bwd`); the free functions below keep the reference's signatures
(rho [grid, spin], grad_rho [grid, spin, 3], grad2rho [grid, spin]) and return [grid].
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .functional import Functional
from .molecule import Molecule

Array = torch.Tensor


def lsda_x_e(rho: Array, clip_cte) -> Array:
    """popular_functionals.py:29-50."""
    return ops.pointwise("LSDA_X", rho, clip=clip_cte)[:, 0]


def b88_x_e(rho: Array, grad_rho: Array, clip_cte: float = 1e-30) -> Array:
    """popular_functionals.py:52-103."""
    return ops.pointwise("B88_X", rho, grad_rho, clip=clip_cte)[:, 0]


def pw92_c_e(rho: Array, clip_cte: float = 1e-30) -> Array:
    """popular_functionals.py:105-139."""
    return ops.pointwise("PW92_C", rho, clip=clip_cte)[:, 0]


def vwn_c_e(rho: Array, clip_cte: float = 1e-30) -> Array:
    """popular_functionals.py:141-195."""
    return ops.pointwise("VWN_C", rho, clip=clip_cte)[:, 0]


def lyp_c_e(rho: Array, grad_rho: Array, grad2rho: Array, clip_cte: float = 1e-30) -> Array:
    """popular_functionals.py:197-269."""
    return ops.pointwise("LYP_C", rho, grad_rho, None, grad2rho, clip=clip_cte)[:, 0]


# ---- feature builders (popular_functionals.py:270-326): one fused kernel per feature set -------------
# the per-point feature set (kernel id) behind each functional's `energy_densities`; `fused_xc_spec` reads the same table
_PW_SET = {"LSDA": "LSDA_X", "B88": "B88_SET", "VWN": "VWN_C", "LYP": "LYP_C", "PW92": "PW92_C", "B3LYP": "B3LYP_SET"}


def lsda_density(molecule: Molecule, clip_cte: float = 1e-30, *_, **__) -> Array:
    return ops.pointwise(_PW_SET["LSDA"], molecule.density(), clip=clip_cte)


def b88_density(molecule: Molecule, clip_cte: float = 1e-30, *_, **__) -> Array:
    return ops.pointwise(_PW_SET["B88"], molecule.density(), molecule.grad_density(), clip=clip_cte)


def vwn_density(molecule: Molecule, clip_cte: float = 1e-30, *_, **__) -> Array:
    return ops.pointwise(_PW_SET["VWN"], molecule.density(), clip=clip_cte)


def pw92_densities(molecule: Molecule, clip_cte: float = 1e-30, *_, **__) -> Array:
    return ops.pointwise(_PW_SET["PW92"], molecule.density(), clip=clip_cte)


def lyp_density(molecule: Molecule, clip_cte: float = 1e-30, *_, **__) -> Array:
    return ops.pointwise(_PW_SET["LYP"], molecule.density(), molecule.grad_density(), None, molecule.lapl_density(), clip=clip_cte)


def b3lyp_exhf_densities(molecule: Molecule, clip_cte: float = 1e-30, *_, **__) -> Array:
    """columns [lsda_x, b88_x, vwn_c, lyp_c] -- popular_functionals.py:306-326."""
    return ops.pointwise(_PW_SET["B3LYP"], molecule.density(), molecule.grad_density(), None, molecule.lapl_density(), clip=clip_cte)


def b3lyp_combine(features: Array, ehf: Array) -> Array:
    """popular_functionals.py:330-338."""
    return torch.cat([features, ehf.sum(dim=(0, 1)).unsqueeze(1)], dim=1)


def _row(values, like: Optional[Array]):
    dev = like.device if isinstance(like, torch.Tensor) else ("cuda" if torch.cuda.is_available() else "cpu")
    return torch.tensor([values], dtype=torch.float64, device=dev)


def b3lyp_coefficients(instance, *args):
    """popular_functionals.py:340-347."""
    a0, ax, ac = 0.2, 0.72, 0.81
    return _ConstRow.get((1 - a0, ax, 1 - ac, ac, a0))


def b3lyp_nograd_densities(molecule: Molecule, *_, **__) -> Array:
    """popular_functionals.py:349-355."""
    return molecule.HF_energy_density([0.0])


def b3lyp_hfgrads(functional, params, molecule: Molecule, ehf, cinputs, densities_wout_hf, omegas=(0.0,)) -> Array:
    """popular_functionals.py:357-372."""
    vxc_hf = molecule.HF_density_grad_2_Fock(functional, params, omegas, ehf, cinputs, densities_wout_hf)
    return vxc_hf.sum(dim=0)


class _ConstRow:
    """Device-resident constant coefficient rows (jnp.array([[...]]) upstream), created once per device."""

    _rows = {}

    @classmethod
    def get(cls, values):
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        key = (values, str(dev))
        if key not in cls._rows:
            cls._rows[key] = torch.tensor([list(values)], dtype=torch.float64, device=dev)
        return cls._rows[key]


_one = lambda self, *_: _ConstRow.get((1.0,))

LSDA = Functional(coefficients=_one, energy_densities=lsda_density, needs=("rho",))
B88 = Functional(coefficients=_one, energy_densities=b88_density, needs=("rho", "grad"))
VWN = Functional(coefficients=_one, energy_densities=vwn_density, needs=("rho",))
LYP = Functional(coefficients=_one, energy_densities=lyp_density, exchange_mask=torch.tensor([]), needs=("rho", "grad", "lapl"))
B3LYP = Functional(
    coefficients=b3lyp_coefficients,
    energy_densities=b3lyp_exhf_densities,
    nograd_densities=b3lyp_nograd_densities,
    densitygrads=b3lyp_hfgrads,
    combine_densities=b3lyp_combine,
    exchange_mask=torch.tensor([1, 1, 0, 0, 1]),
    needs=("rho", "grad", "lapl"),
    needs_omegas=(0.0,),
)
PW92 = Functional(coefficients=_one, energy_densities=pw92_densities, needs=("rho",))


def fused_xc_spec(functional):
    """(pointwise set, constant coefficient row, omegas of the exact-exchange column or None) for the closed-form functionals of
    this module, whose first-order XC build the predictor runs as ONE per-point kernel (ops.xc_point_fused) instead of the
    generic features -> combine -> clip -> quadrature -> autograd chain; None for anything else (user-defined functionals,
    neural functionals).  The rows are the very numbers `coefficients` returns."""
    a0, ax, ac = 0.2, 0.72, 0.81
    table = ((LSDA, _PW_SET["LSDA"], (1.0,), None), (B88, _PW_SET["B88"], (1.0, 1.0), None), (VWN, _PW_SET["VWN"], (1.0,), None),
             (LYP, _PW_SET["LYP"], (1.0,), None), (PW92, _PW_SET["PW92"], (1.0,), None),
             (B3LYP, _PW_SET["B3LYP"], (1 - a0, ax, 1 - ac, ac, a0), (0.0,)))
    for fun, name, row, omegas in table:
        if functional is fun:
            return name, row, omegas
    return None
