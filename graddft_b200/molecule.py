"""`Grid`, `Molecule` and the free functions of grad_dft/molecule.py, backed by the sm_100a kernels.

Mirrors the reference's names, argument meaning and error behaviour (grad_dft/molecule.py:34-336 for the
containers, :341-889 for the free functions) with `torch.Tensor` (float64, CUDA) in place of `jax.Array`.
No arithmetic lives here except n x n housekeeping (make_rdm1, get_occ, orbital_grad: O(n^3) on <= 400^2,
SURVEY.md section 2) -- every grid-sized or n^4-sized contraction is a C-ABI kernel call through `ops`.

Constant per-molecule tensors (ao, grad_ao, grad_n_ao[2], chi) are re-laid-out once into a `PackedBasis`
and reused across SCF iterations / training steps; `Molecule.replace(rdm1=...)` keeps the packed copy.
"""
from __future__ import annotations

import dataclasses
from collections import OrderedDict
from dataclasses import dataclass, fields
from typing import Any, Dict, NamedTuple, Optional, Sequence, Tuple

import torch

from . import ops
from ._lib import GDFT_GRAD, GDFT_HF, GDFT_LAPL, GDFT_RHO, GDFT_TAU

Array = torch.Tensor
F64 = torch.float64


# ---------------------------------------------------------------------------------------------------------
# packed-basis cache for the free functions (the Molecule methods keep their own copy)
# ---------------------------------------------------------------------------------------------------------
def _tkey(t: Optional[Array]):
    return None if t is None else (t.data_ptr(), tuple(t.shape), t._version, str(t.device))


class _BasisCache:
    """Tiny LRU keyed on the identity (pointer, shape, version) of the source tensors."""

    def __init__(self, maxsize: int = 4):
        self.maxsize = maxsize
        self.entries: "OrderedDict[int, Tuple[Dict[str, Any], ops.PackedBasis]]" = OrderedDict()
        self._next = 0

    def get(self, ao, grad_ao=None, grad2_ao=None, chi=None) -> ops.PackedBasis:
        want = {"ao": _tkey(ao), "grad_ao": _tkey(grad_ao), "grad2_ao": _tkey(grad2_ao), "chi": _tkey(chi)}
        for k, (have, basis) in self.entries.items():
            if all(v is None or have[name] == v for name, v in want.items()):
                self.entries.move_to_end(k)
                return basis
        if grad2_ao is not None and grad_ao is None:
            raise ValueError("grad_2_ao needs grad_ao")
        basis = ops.PackedBasis(ao, grad_ao, grad2_ao, chi)
        self.entries[self._next] = (want, basis)
        self._next += 1
        while len(self.entries) > self.maxsize:
            self.entries.popitem(last=False)
        return basis

    def clear(self):
        self.entries.clear()


_CACHE = _BasisCache()


def clear_basis_cache() -> None:
    _CACHE.clear()


def _check(name: str, t: Array, ndim: int) -> None:
    if not isinstance(t, torch.Tensor) or t.dim() != ndim or t.dtype != F64:
        got = f"{tuple(t.shape)} {t.dtype}" if isinstance(t, torch.Tensor) else type(t).__name__
        raise TypeError(f"{name}: expected a float64 tensor with {ndim} dimensions, got {got}")


# ---------------------------------------------------------------------------------------------------------
# free functions  (grad_dft/molecule.py:341-889)
# ---------------------------------------------------------------------------------------------------------
def abs_clip(arr: Array, threshold: float) -> Array:
    """grad_dft/molecule.py:687-689.  One kernel for float64 CUDA tensors (value and VJP); the composite otherwise."""
    if arr.is_cuda and arr.dtype == torch.float64:
        return ops.abs_clip(arr, threshold)
    return torch.where(arr.abs() > threshold, arr, torch.zeros_like(arr))


def density(rdm1: Array, ao: Array, precision=None) -> Array:
    """rho[r,s] -- grad_dft/molecule.py:388-409.  Returns [grid, spin]."""
    _check("rdm1", rdm1, 3), _check("ao", ao, 2)
    return ops.density_forward(_CACHE.get(ao), rdm1, GDFT_RHO)[0]


def grad_density(rdm1: Array, ao: Array, grad_ao: Array, precision=None) -> Array:
    """grad rho[r,s,j] -- grad_dft/molecule.py:414-440.  Returns [grid, spin, 3]."""
    _check("rdm1", rdm1, 3), _check("ao", ao, 2), _check("grad_ao", grad_ao, 3)
    return ops.density_forward(_CACHE.get(ao, grad_ao), rdm1, GDFT_GRAD)[1]


def lapl_density(rdm1: Array, ao: Array, grad_ao: Array, grad_2_ao: Array, precision=None) -> Array:
    """laplacian of rho -- grad_dft/molecule.py:445-474.  Returns [grid, spin]."""
    _check("rdm1", rdm1, 3), _check("ao", ao, 2), _check("grad_ao", grad_ao, 3), _check("grad_2_ao", grad_2_ao, 3)
    return ops.density_forward(_CACHE.get(ao, grad_ao, grad_2_ao), rdm1, GDFT_LAPL)[3]


def kinetic_density(rdm1: Array, grad_ao: Array, precision=None, ao: Optional[Array] = None) -> Array:
    """tau[r,s] -- grad_dft/molecule.py:479-502.  Returns [grid, spin].

    The reference signature has no `ao`; the packed basis always carries plane 0, so when `ao` is not
    given a zero plane of the right shape stands in for it (tau does not read it)."""
    _check("rdm1", rdm1, 3), _check("grad_ao", grad_ao, 3)
    if ao is None:
        for have, basis in _CACHE.entries.values():
            if have["grad_ao"] == _tkey(grad_ao):
                return ops.density_forward(basis, rdm1, GDFT_TAU)[2]
        ao = torch.zeros(grad_ao.shape[:2], dtype=F64, device=grad_ao.device)
    return ops.density_forward(_CACHE.get(ao, grad_ao), rdm1, GDFT_TAU)[2]


def HF_energy_density(rdm1: Array, ao: Array, chi: Array, precision=None) -> Array:
    """e_HF[w,s,r] -- grad_dft/molecule.py:507-541.  Returns [omega, spin, grid]."""
    _check("rdm1", rdm1, 3), _check("ao", ao, 2), _check("chi", chi, 4)
    return ops.density_forward(_CACHE.get(ao, chi=chi), rdm1, GDFT_HF)[4]


def HF_density_grad_2_Fock(grid, functional, params, chi: Array, ao: Array, ehf: Array, coefficient_inputs,
                           densities_wout_hf: Array, chunk_size=None, precision=None, _basis=None) -> Array:
    """grad_dft/molecule.py:545-613: g = dE_xc/d e_HF through the densities, then F[w,s] = -1/2 ao^T diag(g) chi.
    `chunk_size` is accepted and ignored (the GEMM kernel needs no chunking).  Returns [omega, spin, n, n]."""
    ehf_leaf = ehf.detach().requires_grad_(True)
    higher = torch.is_grad_enabled()  # under a differentiable SCF loop g itself depends on params / the features
    with torch.enable_grad():
        densities = functional.combine_densities(densities_wout_hf, ehf_leaf)
        e = functional.xc_energy(params, grid, coefficient_inputs, densities)
    (gr,) = torch.autograd.grad(e, ehf_leaf, create_graph=higher)
    basis = _basis if _basis is not None else _CACHE.get(ao, chi=chi)
    return ops.hf_fock(basis, gr)


def HF_coefficient_input_grad_2_Fock(grid, functional, params, chi: Array, ao: Array, ehf: Array, cinputs_wout_hf,
                                     densities: Array, chunk_size=None, precision=None, _basis=None) -> Array:
    """grad_dft/molecule.py:617-685: same as above with the derivative taken through the coefficient inputs."""
    ehf_leaf = ehf.detach().requires_grad_(True)
    higher = torch.is_grad_enabled()
    with torch.enable_grad():
        cinputs = functional.combine_inputs(cinputs_wout_hf, ehf_leaf)
        e = functional.xc_energy(params, grid, cinputs, densities)
    (gr,) = torch.autograd.grad(e, ehf_leaf, create_graph=higher)
    basis = _basis if _basis is not None else _CACHE.get(ao, chi=chi)
    return ops.hf_fock(basis, gr)


def coulomb_potential(rdm1: Array, rep_tensor: Array, precision=None) -> Array:
    """J[p,q] = sum_rt (pq|rt) P[r,t] -- grad_dft/molecule.py:788-811 (rdm1 is the spin-summed [n,n] matrix)."""
    _check("rdm1", rdm1, 2), _check("rep_tensor", rep_tensor, 4)
    return ops.coulomb_j_auto(rdm1, rep_tensor)


def coulomb_energy(rdm1: Array, rep_tensor: Array, precision=None) -> Array:
    """E_J = 1/2 <P, J> -- grad_dft/molecule.py:763-783."""
    v = coulomb_potential(rdm1, rep_tensor)
    return (rdm1 * v).sum() / 2.0


def one_body_energy(rdm1: Array, h1e: Array, precision=None) -> Array:
    """grad_dft/molecule.py:738-757."""
    return (rdm1 * h1e).sum()


def nonXC(rdm1: Array, h1e: Array, rep_tensor: Array, nuclear_repulsion, precision=None) -> Array:
    """E_nuc + E_1 + E_J -- grad_dft/molecule.py:697-733."""
    return nuclear_repulsion + one_body_energy(rdm1, h1e) + coulomb_energy(rdm1, rep_tensor)


def make_rdm1(mo_coeff: Array, mo_occ: Array, precision=None) -> Array:
    """D[s,i,k] = sum_j C[s,i,j] occ[s,j] C[s,k,j] -- grad_dft/molecule.py:815-846 (n x n, host framework)."""
    return torch.einsum("sij,sj,skj->sik", mo_coeff, mo_occ, mo_coeff)


def get_occ(mo_energies: Array, nelecs: Array, naos: int) -> Array:
    """Aufbau occupations -- grad_dft/molecule.py:851-889: the nelecs[s] lowest orbitals get 1 (stable argsort)."""
    idx = torch.argsort(mo_energies, dim=1, stable=True)
    rank = torch.empty_like(idx)
    ar = torch.arange(naos, device=mo_energies.device).expand_as(idx)
    rank.scatter_(1, idx, ar)
    return (rank < nelecs.to(mo_energies.device).reshape(2, 1)).to(mo_energies.dtype)


def orbital_grad(mo_coeff: Array, mo_occ: Array, F: Array, precision=None) -> Array:
    """C_vir^T F C_occ summed over spin with zero-masked blocks -- grad_dft/molecule.py:344-381."""
    occ = (mo_occ > 0).unsqueeze(1)
    vir = (mo_occ == 0).unsqueeze(1)
    zero = torch.zeros_like(mo_coeff)
    return torch.einsum("sab,sac,scd->bd", torch.where(vir, mo_coeff, zero), F, torch.where(occ, mo_coeff, zero))


# ---------------------------------------------------------------------------------------------------------
# containers  (grad_dft/molecule.py:34-336, 895-953)
# ---------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Grid:
    """grad_dft/molecule.py:34-69."""

    coords: Array
    weights: Array

    def __len__(self):
        return self.weights.shape[0]

    def to_dict(self) -> dict:
        return {"coords": self.coords, "weights": self.weights}

    def integrate(self, vals: Array, axis: int = 0) -> Array:
        return torch.tensordot(self.weights, vals, dims=([0], [axis]))

    def replace(self, **kw) -> "Grid":
        return dataclasses.replace(self, **kw)


def _okey(omegas) -> tuple:
    return tuple(float(o) for o in (omegas.tolist() if isinstance(omegas, torch.Tensor) else omegas))


_NAMES = {"rho": GDFT_RHO, "grad": GDFT_GRAD, "tau": GDFT_TAU, "lapl": GDFT_LAPL}
_SLOT = {GDFT_RHO: 0, GDFT_GRAD: 1, GDFT_TAU: 2, GDFT_LAPL: 3}
_BASIS_FIELDS = ("ao", "grad_ao", "grad_n_ao", "chi")


@dataclass(frozen=True)
class Molecule:
    """grad_dft/molecule.py:72-336: same fields, same method names.  `replace` is flax.struct's."""

    grid: Grid
    atom_index: Any
    nuclear_pos: Any
    ao: Array
    grad_ao: Array
    grad_n_ao: Any
    rdm1: Array
    nuclear_repulsion: Any
    h1e: Array
    vj: Any
    mo_coeff: Array
    mo_occ: Array
    mo_energy: Array
    mf_energy: Any = None
    s1e: Optional[Array] = None
    omegas: Any = None
    chi: Optional[Array] = None
    rep_tensor: Optional[Array] = None
    energy: Any = None
    basis: Any = None
    name: Any = None
    spin: Any = 0
    charge: Any = 0
    unit_Angstrom: Any = True
    grid_level: Any = 2
    scf_iteration: Any = 50
    fock: Optional[Array] = None

    # ---- plumbing ---------------------------------------------------------------------------------------
    def replace(self, **kw) -> "Molecule":
        new = dataclasses.replace(self, **kw)
        if not any(k in kw for k in _BASIS_FIELDS):
            for k in ("_packed", "_omega_list"):
                if k in self.__dict__:
                    object.__setattr__(new, k, self.__dict__[k])
        if "_shard" in self.__dict__:
            object.__setattr__(new, "_shard", self.__dict__["_shard"])
        return new

    @property
    def grid_size(self):
        return len(self.grid)

    @property
    def packed_basis(self) -> ops.PackedBasis:
        """The planar copy of ao / grad_ao / grad_n_ao[2] / chi, built on first use and shared by `replace`."""
        pb = self.__dict__.get("_packed")
        if pb is None:
            g2 = None
            if self.grad_n_ao is not None:
                try:
                    g2 = self.grad_n_ao[2]
                except (KeyError, IndexError, TypeError):
                    g2 = None
            pb = ops.PackedBasis(self.ao, self.grad_ao, g2 if self.grad_ao is not None else None, self.chi)
            object.__setattr__(self, "_packed", pb)
        return pb

    def _memo(self) -> dict:
        key = (self.rdm1.data_ptr(), self.rdm1._version, id(self.rdm1))
        m = self.__dict__.get("_memo_store")
        if m is None or m.get("key") != key:
            m = {"key": key}
            object.__setattr__(self, "_memo_store", m)
        return m

    def prefetch(self, *names: str, omegas: Optional[Sequence[float]] = None) -> None:
        """Compute several grid quantities of the current rdm1 in ONE kernel launch (they share T = ao D)
        and keep them for the method calls that follow.  names from {"rho","grad","tau","lapl"}."""
        flags = 0
        for nm in names:
            flags |= _NAMES[nm]
        basis = self.packed_basis
        if omegas is not None and len(omegas):
            basis = basis.select_chi(self._omega_indices(omegas))
            flags |= GDFT_HF
        outs = ops.density_forward(basis, self.rdm1, flags)
        m = self._memo()
        for f, slot in _SLOT.items():
            if flags & f:
                m[f] = outs[slot]
        if flags & GDFT_HF:
            m[("hf", _okey(omegas))] = outs[4]

    def _quantity(self, flag: int) -> Array:
        m = self._memo()
        if flag not in m:
            pb = self.packed_basis
            if flag in (GDFT_RHO, GDFT_GRAD):
                names = ["rho"] + (["grad"] if pb.nplanes >= 4 else [])
            else:
                names = ["rho", "grad", "tau"] + (["lapl"] if pb.nplanes >= 5 else [])
            if flag == GDFT_LAPL and pb.nplanes < 5:
                raise ValueError("lapl_density needs grad_n_ao[2]")
            if flag in (GDFT_GRAD, GDFT_TAU) and pb.nplanes < 4:
                raise ValueError("grad_ao has not been loaded")
            self.prefetch(*names)
        return m[flag]

    def to_dict(self) -> dict:
        grid_dict = self.grid.to_dict()
        rest = {f.name: getattr(self, f.name) for f in fields(self)[1:]}
        return dict(**grid_dict, **rest)

    # ---- grid quantities ------------------------------------------------------------------------------
    def density(self, *args, **kwargs) -> Array:
        return self._quantity(GDFT_RHO)

    def grad_density(self, *args, **kwargs) -> Array:
        return self._quantity(GDFT_GRAD)

    def lapl_density(self, *args, **kwargs) -> Array:
        return self._quantity(GDFT_LAPL)

    def kinetic_density(self, *args, **kwargs) -> Array:
        return self._quantity(GDFT_TAU)

    def _omega_indices(self, omegas) -> list:
        if self.chi is None:
            raise ValueError("Precomputed chi tensor has not been loaded.")
        have = self.__dict__.get("_omega_list")
        if have is None:
            have = [float(o) for o in (self.omegas.tolist() if isinstance(self.omegas, torch.Tensor) else self.omegas)]
            object.__setattr__(self, "_omega_list", have)
        want = [float(o) for o in (omegas.tolist() if isinstance(omegas, torch.Tensor) else omegas)]
        for o in want:
            if o not in have:
                raise ValueError(f"The molecule.chi tensor does not contain omega value {o}, only {self.omegas}")
        return [have.index(o) for o in want]

    def select_HF_omegas(self, omegas) -> Array:
        indices = self._omega_indices(omegas)
        return torch.stack([self.chi[:, i] for i in indices], dim=1)

    def HF_energy_density(self, omegas, *args, **kwargs) -> Array:
        m = self._memo()
        key = ("hf", _okey(omegas))
        if key not in m:
            basis = self.packed_basis.select_chi(self._omega_indices(omegas))
            m[key] = ops.density_forward(basis, self.rdm1, GDFT_HF)[4]
        return m[key]

    def hf_fock_summed(self, omegas, g: Array) -> Array:
        """sum over omega of -1/2 ao^T diag(g[w,s]) chi[w,s] ([2,n,n]) for a cotangent g[len(omegas), 2, N] that is already at hand
        (the predictor's merged exact-exchange routes): one GEMM of two units for up to two omegas."""
        return ops.hf_fock_sum(self.packed_basis.select_chi(self._omega_indices(omegas)), g)

    def HF_density_grad_2_Fock(self, functional, params, omegas, ehf, coefficient_inputs, densities_wout_hf, **kwargs) -> Array:
        basis = self.packed_basis.select_chi(self._omega_indices(omegas))
        b = self._memo().get("xc_build")
        if (b is not None and not torch.is_grad_enabled() and b.matches(functional, params) and ehf is b.nograd_densities
                and coefficient_inputs is b.cinputs and densities_wout_hf is b.grad_densities):
            # the coefficients of this very (params, coefficient_inputs) pair were evaluated by the XC build of the same
            # predictor call: molecule.py:600-604 with them as constants (no second pass through the network)
            if b.g_densities is not None:
                # ... and dE_xc/d e_HF through the densities is the cotangent that reached the stop_gradient boundary in that
                # build's own backward pass (its densities are abs_clip'ed, a difference of at most 1e-30 in magnitude --
                # DESIGN.md section 4): no second quadrature pass either
                return ops.hf_fock(basis, b.g_densities)
            ehf_leaf = ehf.detach().requires_grad_(True)
            with torch.enable_grad():
                e = ops.xc_integrate(b.coefficients, functional.combine_densities(densities_wout_hf, ehf_leaf), self.grid.weights, 1e-30)
            (gr,) = torch.autograd.grad(e, ehf_leaf)
            return ops.hf_fock(basis, gr)
        return HF_density_grad_2_Fock(self.grid, functional, params, None, self.ao, ehf, coefficient_inputs, densities_wout_hf,
                                      _basis=basis, **kwargs)

    def HF_coefficient_input_grad_2_Fock(self, functional, params, omegas, ehf, cinputs_wout_hf, densities, **kwargs) -> Array:
        basis = self.packed_basis.select_chi(self._omega_indices(omegas))
        b = self._memo().get("xc_build")
        if (b is not None and not torch.is_grad_enabled() and b.matches(functional, params) and b.g_cinputs is not None
                and ehf is b.nograd_cinputs and cinputs_wout_hf is b.grad_cinputs and densities is b.densities_raw):
            # molecule.py:672-676 asks for dE_xc/d e_HF through the coefficient inputs: the cotangent that reached the
            # stop_gradient boundary in the XC build's own backward pass (same network, same inputs; the densities there
            # are abs_clip'ed, a difference of at most 1e-30 in magnitude -- DESIGN.md section 4)
            return ops.hf_fock(basis, b.g_cinputs)
        return HF_coefficient_input_grad_2_Fock(self.grid, functional, params, None, self.ao, ehf, cinputs_wout_hf, densities,
                                                _basis=basis, **kwargs)

    # ---- n x n ------------------------------------------------------------------------------------------
    def get_coulomb_potential(self, *args, **kwargs) -> Array:
        return coulomb_potential(self.rdm1.sum(dim=0), self.rep_tensor)

    def nonXC(self, *args, **kwargs) -> Array:
        return nonXC(self.rdm1.sum(dim=0), self.h1e, self.rep_tensor, self.nuclear_repulsion)

    def make_rdm1(self) -> Array:
        return make_rdm1(self.mo_coeff, self.mo_occ)

    def get_occ(self) -> Array:
        if (self.mo_energy.is_cuda and self.mo_energy.dim() == 2 and self.mo_energy.shape[0] == 2 and self.mo_energy.dtype == torch.float64
                and self.mo_occ.shape == self.mo_energy.shape):
            return ops.aufbau_occupations(self.mo_energy, self.mo_occ)  # same ranks as the argsort below, one kernel
        nelecs = self.mo_occ.sum(dim=1).round().to(torch.int64)
        return get_occ(self.mo_energy, nelecs, self.mo_occ.shape[1])

    def get_mo_grads(self, *args, **kwargs) -> Array:
        return orbital_grad(self.mo_coeff, self.mo_occ, self.fock)


class Reaction(NamedTuple):
    """grad_dft/molecule.py:895-912."""

    reactants: Sequence[Molecule]
    products: Sequence[Molecule]
    reactant_numbers: Sequence[int]
    product_numbers: Sequence[int]
    energy: float
    name: Any = None


def molecule_from_tensors(mol: Dict[str, Array], device=None) -> Molecule:
    """Build a `Molecule` from a dict keyed by the reference's field names (as `synthetic.synthetic_molecule`
    or an HDF5 loader would give), moving tensors to `device`.  `grad_n_ao2` stands for grad_n_ao[2]."""

    def mv(t):
        return t.to(device) if (device is not None and isinstance(t, torch.Tensor)) else t

    g2 = mol.get("grad_n_ao2")
    return Molecule(
        grid=Grid(mv(mol.get("coords")), mv(mol["weights"])), atom_index=mol.get("atom_index"), nuclear_pos=mol.get("nuclear_pos"),
        ao=mv(mol["ao"]), grad_ao=mv(mol.get("grad_ao")), grad_n_ao=({2: mv(g2)} if g2 is not None else None), rdm1=mv(mol["rdm1"]),
        nuclear_repulsion=mv(mol.get("nuclear_repulsion")), h1e=mv(mol.get("h1e")), vj=mv(mol.get("vj")), mo_coeff=mv(mol.get("mo_coeff")),
        mo_occ=mv(mol.get("mo_occ")), mo_energy=mv(mol.get("mo_energy")), s1e=mv(mol.get("s1e")), omegas=mol.get("omegas"),
        chi=mv(mol.get("chi")), rep_tensor=mv(mol.get("rep_tensor")), fock=mv(mol.get("fock")),
    )
