"""`energy_predictor` -- grad_dft/train.py:34-218, re-stated over the kernel-backed ops.

`predict(params, molecule) -> (energy, fock)` follows the reference step by step: value-and-grad of the XC
energy with respect to rdm1 (train.py:86-121,147), h1e + J (148), E = Exc + nonXC (150), clip / symmetrise /
clip (161-163), the explicit exact-exchange Fock terms of hybrids with `fock += V + V^T` and a clip after each
(200-213), final clip (215).  J and E_J come from one ERI sweep (the reference sweeps twice; under jit XLA
merges them -- SURVEY.md Appendix B).
"""
from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import ops
from .functional import Functional, stop_gradient
from .molecule import Molecule, abs_clip

Array = torch.Tensor


def _requires_grad(params) -> bool:
    if isinstance(params, torch.Tensor):
        return params.requires_grad
    if isinstance(params, dict):
        return any(_requires_grad(v) for v in params.values())
    if isinstance(params, (list, tuple)):
        return any(_requires_grad(v) for v in params)
    return False


def xc_energy_and_grads(functional: Functional, params, rdm1: Array, atoms: Molecule, *args, create_graph: Optional[bool] = None,
                        **functional_kwargs) -> Tuple[Array, Array, Molecule]:
    """value_and_grad(argnums=1) of the XC energy (grad_dft/train.py:86-121): one "XC build" = forward
    (densities -> features -> E_xc) + VJP (-> V_xc [2,n,n], un-symmetrised).  Also returns the molecule
    carrying the differentiated rdm1 (its cached grid quantities are reused by the hybrid terms)."""
    keep_exc_graph = torch.is_grad_enabled() and (_requires_grad(params) or rdm1.requires_grad)
    if not keep_exc_graph and not create_graph and not args and not functional_kwargs and rdm1.is_cuda and os.environ.get("GDFT_FUSED_XC", "1") != "0":
        from .popular_functionals import fused_xc_spec
        spec = fused_xc_spec(functional)
        if spec is not None:
            return _fused_xc_build(functional, spec, params, rdm1, atoms)
    if create_graph is None:
        # V_xc itself differentiable (w.r.t. params and rdm1) only when the density matrix already carries a graph,
        # i.e. inside a differentiable SCF loop (evaluate.py:917-1038); energy-only losses need first order only
        create_graph = torch.is_grad_enabled() and rdm1.requires_grad
    leaf = rdm1 if (create_graph and rdm1.requires_grad) else rdm1.detach().requires_grad_(True)
    tap = (not create_graph) and bool(functional.nograd_densities or functional.nograd_coefficient_inputs)
    with torch.enable_grad(), ops.first_order_build(not create_graph):
        at = atoms.replace(rdm1=leaf)
        if tap:
            # hybrids: the explicit exact-exchange routes (train.py:200-213) ask for dE_xc/d e_HF through the densities and
            # through the coefficient inputs; both cotangents arrive at the stop_gradient boundary of THIS backward pass,
            # and the coefficients they need are the ones evaluated here.  Keep them with the build (under jit XLA's CSE
            # merges these re-evaluations; here the predictor does, see Molecule.HF_*_grad_2_Fock).
            clip = functional_kwargs.get("clip_cte", args[0] if args else 1e-30)
            ft = functional._features(at, True, *args, **functional_kwargs)
            xkw = {k: v for k, v in functional_kwargs.items() if k != "clip_cte"}
            coefficients = functional.coefficients_for(params, ft["cinputs"], ft["densities"], **xkw)
            exc = ops.xc_integrate(coefficients, ft["densities"], at.grid.weights, clip)
            taps = [t for t in (ft["tap_d"], ft["tap_c"]) if t is not None]
            grads = torch.autograd.grad(exc, [leaf] + taps, retain_graph=keep_exc_graph, allow_unused=True)
            fock_xc = grads[0]
            gt = dict(zip([k for k in ("tap_d", "tap_c") if ft[k] is not None], grads[1:]))

            def const(x):
                return x.detach() if isinstance(x, torch.Tensor) else x

            at._memo()["xc_build"] = XCBuild(
                functional=functional, params_key=_params_key(params), clip=float(clip), coefficients=const(coefficients),
                grad_densities=const(ft["grad_densities"]), nograd_densities=const(ft["tap_d"]), densities_raw=const(ft["densities_raw"]),
                grad_cinputs=const(ft["grad_cinputs"]), nograd_cinputs=const(ft["tap_c"]), cinputs=const(ft["cinputs"]),
                g_densities=gt.get("tap_d"), g_cinputs=gt.get("tap_c"))
        else:
            densities = functional.compute_densities(at, *args, **functional_kwargs)
            cinputs = functional.compute_coefficient_inputs(at, *args)
            exc = functional.xc_energy(params, at.grid, cinputs, densities, **functional_kwargs)
            (fock_xc,) = torch.autograd.grad(exc, leaf, create_graph=create_graph, retain_graph=keep_exc_graph or create_graph)
    if not (keep_exc_graph or create_graph):
        exc = exc.detach()  # otherwise E_xc stays differentiable w.r.t. params (first order: energy losses, train.py:312-359)
    return exc, fock_xc, at


def _fused_xc_build(functional, spec, params, rdm1: Array, atoms: Molecule):
    """The first-order XC build of a closed-form functional as K1 -> ONE per-point kernel -> K2 (popular_functionals.fused_xc_spec,
    ops.xc_point_fused): the same arithmetic as the generic chain of xc_energy_and_grads, evaluated once per point.  Leaves the
    cotangent at the exact-exchange stop_gradient boundary on the molecule for the explicit Fock term, like the generic path."""
    from ._lib import GDFT_GRAD, GDFT_HF, GDFT_LAPL, GDFT_RHO

    name, row, omegas = spec
    at = atoms.replace(rdm1=rdm1.detach())
    want_grad, want_lapl = "grad" in functional.needs, "lapl" in functional.needs  # what the functional's own feature set reads
    flags = GDFT_RHO | (GDFT_GRAD if want_grad else 0) | (GDFT_LAPL if want_lapl else 0) | (GDFT_HF if omegas else 0)
    basis = at.packed_basis.select_chi(at._omega_indices(omegas)) if omegas else at.packed_basis
    with torch.no_grad():
        rho, grho, _, lapl, ehf = ops._density_fwd_raw(basis, at.rdm1, flags)
        exc, rb, gb, _, lb, eb = ops.xc_point_fused(name, 1e-30, row, rho, grho, None, lapl, ehf, at.grid.weights)
        fock_xc = ops.density_transpose(basis, rb, gb, None, lb)
    at._memo()["xc_build"] = XCBuild(
        functional=functional, params_key=_params_key(params), clip=1e-30, coefficients=None, grad_densities=None, nograd_densities=None,
        densities_raw=None, grad_cinputs=None, nograd_cinputs=None, cinputs=None, g_densities=eb, g_cinputs=None)
    return exc, fock_xc, at


def _params_key(params):
    return tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in _leaves(params)) if params is not None else None


class XCBuild:
    """What one first-order XC build (forward + VJP) of a hybrid functional leaves on the molecule for the explicit
    exact-exchange routes of the same predictor call: the feature tensors (as constants), the coefficients, and the
    cotangents of E_xc at the two stop_gradient boundaries.  Valid only for the very tensors it holds (`is` checks)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def matches(self, functional, params) -> bool:
        return self.functional is functional and self.clip == 1e-30 and self.params_key == _params_key(params)


def energy_predictor(functional: Functional, nlc_functional=None, clip_cte: float = 1e-30, differentiable_fock: Optional[bool] = None,
                     **kwargs) -> Callable:
    """grad_dft/train.py:34-218.  `differentiable_fock` (not upstream, where jax traces everything): True keeps the
    returned Fock matrix differentiable w.r.t. params and rdm1 (second-order kernels; what `diff_scf_loop` asks for
    when params require grad), False never does, None = only when rdm1 already carries a graph."""
    if nlc_functional is not None:
        raise NotImplementedError("nlc_functional: the reference path raises NameError here (train.py:117-120)")

    def explicit_terms(params, at: Molecule, differentiable: bool, *args):
        """train.py:165-213: the explicit exact-exchange Fock terms V (one per HF route of the functional); the
        features they need are cached on `at`, so nothing is recomputed.  They do not depend on the Fock matrix
        being assembled, only on (params, rdm1)."""
        build = at._memo().get("xc_build") if not differentiable else None
        if build is not None and build.matches(functional, params):
            # same rdm1, same params: the features are the ones the XC build just evaluated (pure functions of them)
            from .functional import standard_hf_routes
            omegas = standard_hf_routes(functional)
            gs = [g for g, route in ((build.g_densities, functional.densitygrads), (build.g_cinputs, functional.coefficient_input_grads))
                  if route and g is not None]
            n_routes = int(bool(functional.densitygrads)) + int(bool(functional.coefficient_input_grads))
            if omegas is not None and gs and len(gs) == n_routes and at.rdm1.is_cuda and all(tuple(g.shape) == tuple(gs[0].shape) for g in gs):
                # built-in routes: both cotangents dE_xc/d e_HF are by-products of the build's backward pass and V is linear in
                # them -- one GEMM for all routes and omegas (functional.standard_hf_routes)
                with torch.no_grad():
                    return [at.hf_fock_summed(omegas, gs[0] if len(gs) == 1 else gs[0] + gs[1])]
            terms = []
            with torch.no_grad():
                if functional.densitygrads:
                    terms.append(functional.densitygrads(functional, params, at, build.nograd_densities, build.cinputs, build.grad_densities))
                if functional.coefficient_input_grads:
                    terms.append(functional.coefficient_input_grads(functional, params, at, build.nograd_cinputs, build.grad_cinputs,
                                                                    build.densities_raw))
            return terms
        with torch.set_grad_enabled(differentiable):
            if functional.energy_densities and functional.densitygrads:
                grad_densities = functional.energy_densities(at, *args, **kwargs)
                nograd_densities = stop_gradient(functional.nograd_densities(at, *args, **kwargs))
                densities = functional.combine_densities(grad_densities, nograd_densities)
            elif functional.energy_densities:
                grad_densities, nograd_densities = functional.energy_densities(at, *args, **kwargs), None
                densities = grad_densities
            elif functional.densitygrads:
                grad_densities, nograd_densities = None, stop_gradient(functional.nograd_densities(at, *args, **kwargs))
                densities = nograd_densities
            else:
                densities, grad_densities, nograd_densities = None, None, None

            if functional.coefficient_input_grads and functional.coefficient_inputs:
                grad_cinputs = functional.coefficient_inputs(at, *args, **kwargs)
                nograd_cinputs = stop_gradient(functional.nograd_coefficient_inputs(at, *args, **kwargs))
                cinputs = functional.combine_inputs(grad_cinputs, nograd_cinputs)
            elif functional.coefficient_inputs:
                grad_cinputs, nograd_cinputs = functional.coefficient_inputs(at, *args, **kwargs), None
                cinputs = grad_cinputs
            elif functional.coefficient_input_grads:
                grad_cinputs, nograd_cinputs = None, stop_gradient(functional.nograd_coefficient_inputs(at, *args, **kwargs))
                cinputs = nograd_cinputs
            else:
                cinputs, grad_cinputs, nograd_cinputs = None, None, None

            # first order: the features enter the explicit terms as constants; differentiable Fock: they keep their
            # graph, as under jax.grad of the whole loop (train.py:200-213 passes them un-stopped)
            def through(t):
                return t if (differentiable or not isinstance(t, torch.Tensor)) else t.detach()

            terms = []
            if functional.densitygrads:
                terms.append(functional.densitygrads(functional, params, at, nograd_densities, through(cinputs), through(grad_densities)))
            if functional.coefficient_input_grads:
                terms.append(functional.coefficient_input_grads(functional, params, at, nograd_cinputs, through(grad_cinputs), through(densities)))
        return terms

    def predict(params, atoms: Molecule, *args) -> Tuple[Array, Array]:
        shard = atoms.__dict__.get("_shard")
        create_graph = differentiable_fock
        if differentiable_fock and not (torch.is_grad_enabled() and (_requires_grad(params) or atoms.rdm1.requires_grad)):
            create_graph = False
        into = None
        if shard is not None and atoms.rdm1.is_cuda:
            from . import distributed as gdist
            n2 = atoms.rdm1.shape[-1] ** 2
            nv = int(bool(functional.densitygrads)) + int(bool(functional.coefficient_input_grads))
            # payload [V_xc | E_xc | J | V_HF...]: the density VJP's second-stage reduce writes V_xc straight into it
            into = gdist.payload_segment(0, [2 * n2, 1] + ([n2] if shard.eri_sharded else []) + [2 * n2] * nv, atoms.rdm1.device, shard.group)
        with ops.density_bwd_into(into):
            exc, fock_xc, at = xc_energy_and_grads(functional, params, atoms.rdm1, atoms, *args, create_graph=create_graph)
        differentiable = fock_xc.requires_grad
        P = atoms.rdm1.sum(dim=0)
        if shard is not None:
            # grid rows (and optionally the (p,q) rows of rep_tensor) live on `world` GPUs: every grid-derived
            # quantity below is a partial sum; ONE all-reduce of [E_xc | V_xc | J | V_HF...] closes the build
            if differentiable or atoms.rdm1.requires_grad:
                raise NotImplementedError("the grid-sharded predictor is first-order in params only (no create_graph through the Fock matrix)")
            from . import distributed as gdist
            vterms = explicit_terms(params, at, False, *args)
            J = gdist.local_coulomb(P, atoms.rep_tensor, shard)
            (fock_xc, exc_sum, J, *vterms) = gdist.allreduce_sum_packed(
                [fock_xc, exc.detach().reshape(1), J, *vterms], group=shard.group, skip=() if shard.eri_sharded else (2,))
            exc_sum = exc_sum.reshape(())
            exc = exc + (exc_sum - exc.detach()) if exc.requires_grad else exc_sum  # value: the global sum; gradient: identity
        elif differentiable or atoms.rdm1.requires_grad:
            J = ops.coulomb_j_auto(P, atoms.rep_tensor)
        else:
            J = ops.coulomb_j_auto(P.detach(), atoms.rep_tensor)
        enuc = atoms.nuclear_repulsion
        if P.is_cuda and isinstance(enuc, torch.Tensor) and enuc.is_cuda and not (J.requires_grad or P.requires_grad or enuc.requires_grad):
            energy = exc + ops.nonxc_energy(P, atoms.h1e, J, atoms.nuclear_repulsion)  # train.py:150, molecule.py:727-733 in one kernel
        else:
            energy = exc + (atoms.nuclear_repulsion + (P * atoms.h1e).sum() + (P * J).sum() / 2.0)  # train.py:150, molecule.py:727-733

        if fock_xc.requires_grad or J.requires_grad:
            fock = atoms.h1e + J + fock_xc
            fock = abs_clip(fock, clip_cte)
            fock = 0.5 * (fock + fock.transpose(1, 2))
            fock = abs_clip(fock, clip_cte)
        else:
            fock = ops.fock_assemble(atoms.h1e, J, fock_xc, clip_cte)  # train.py:148-163 in one kernel

        if shard is None:
            vterms = explicit_terms(params, at, differentiable, *args)
        for vxc_expl in vterms:  # train.py:200-213: fock += V + V^T, clip after each
            fock = _add_sym(fock, vxc_expl, clip_cte)
        fock = abs_clip(fock, clip_cte)
        return energy, fock

    def energy_only(params, atoms: Molecule, *args) -> Array:
        """The energy output alone: E_xc forward + nonXC, no VJP.  Under `jit` XLA drops the unused Fock matrix of an
        energy-only loss as dead code (train.py:480-575 only read `.energy`); the losses here ask for this entry instead."""
        if atoms.__dict__.get("_shard") is not None:
            return predict(params, atoms, *args)[0]
        with ops.first_order_build():
            exc = functional.energy_xc_only(params, atoms, *args, **kwargs)
        P = atoms.rdm1.sum(dim=0)
        if atoms.rdm1.requires_grad and torch.is_grad_enabled():
            EJ = (P * ops.coulomb_j_auto(P, atoms.rep_tensor)).sum() / 2.0
        else:
            EJ = ops.coulomb_j_and_energy(P, atoms.rep_tensor)[1]
        return exc + (atoms.nuclear_repulsion + (P * atoms.h1e).sum() + EJ)

    def energy_only_batch(params, atoms_list: Sequence[Molecule], *args, max_points: int = 600_000) -> List[Array]:
        """`energy_only` for several molecules with ONE pass of the coefficient network per group of molecules (their
        grid rows concatenated, at most `max_points` per group).  The reference evaluates a batch serially
        (train.py:519-528); the network acts row by row, so concatenating rows changes nothing but the number of launches:
        a training step over 64 small molecules is otherwise bound by the host's launch rate (1472 GEMM launches, the
        host enqueueing for 240 of 267 ms).  Features, quadrature and the non-XC terms stay per molecule."""
        atoms_list = list(atoms_list)
        if (functional.coefficient_inputs is None and functional.nograd_coefficient_inputs is None) or any(
                a.__dict__.get("_shard") is not None for a in atoms_list):
            return [energy_only(params, a, *args) for a in atoms_list]
        out: List[Optional[Array]] = [None] * len(atoms_list)
        start = 0
        while start < len(atoms_list):
            end, pts = start, 0
            while end < len(atoms_list) and (end == start or pts + atoms_list[end].grid_size <= max_points):
                pts += atoms_list[end].grid_size
                end += 1
            group = atoms_list[start:end]
            with ops.first_order_build():
                dens = [functional.compute_densities(a, *args, **kwargs) for a in group]
                cins = [functional.compute_coefficient_inputs(a, *args) for a in group]
                like = torch.empty((pts, dens[0].shape[1]), dtype=dens[0].dtype, device="meta")
                coeffs = functional.coefficients_for(params, torch.cat(cins, dim=0), like, **kwargs)
                if coeffs.shape[0] != pts:  # row-independent coefficients: nothing to batch
                    return [energy_only(params, a, *args) for a in atoms_list]
                for k, (a, d, c) in enumerate(zip(group, dens, coeffs.split([a.grid_size for a in group], dim=0))):
                    exc = ops.xc_integrate(c, d, a.grid.weights, clip_cte)
                    P = a.rdm1.sum(dim=0)
                    if a.rdm1.requires_grad and torch.is_grad_enabled():
                        EJ = (P * ops.coulomb_j_auto(P, a.rep_tensor)).sum() / 2.0
                    else:
                        EJ = ops.coulomb_j_and_energy(P, a.rep_tensor)[1]
                    out[start + k] = exc + (a.nuclear_repulsion + (P * a.h1e).sum() + EJ)
            start = end
        return out

    predict.energy_only = energy_only
    predict.energy_only_batch = energy_only_batch
    return predict


def _add_sym(fock: Array, v: Array, clip_cte: float) -> Array:
    """fock += V + V^T; abs_clip   (train.py:205-206, 212-213)."""
    if fock.requires_grad or v.requires_grad:
        return abs_clip(fock + (v + v.transpose(1, 2)), clip_cte)
    return ops.fock_add_sym_(fock, v, clip_cte)


def Harris_energy_predictor(functional: Functional, **kwargs) -> Callable:
    """grad_dft/train.py:220-308: E_Harris = sum_i occ_i eps_i - E_J[P] + E_xc - <rdm1, V_xc> + E_nuc, with V_xc the
    (un-symmetrised) derivative of E_xc w.r.t. rdm1 from the same fused forward + VJP build."""

    def Harris_energy(params, molecule: Molecule, *args, **hkwargs) -> Array:
        energy = (molecule.mo_occ * molecule.mo_energy).sum()
        P = molecule.rdm1.sum(dim=0)
        coulomb_e = -(P * ops.coulomb_j(P, molecule.rep_tensor)).sum() / 2.0
        # jax.grad of the Harris energy differentiates THROUGH xcfock (the -<rdm1, dV_xc/dtheta> term), so V_xc must keep
        # its graph whenever anything upstream of it is being differentiated; the first-order tap path of hybrids would
        # hand back a constant V_xc and is not used then
        if "create_graph" not in hkwargs:
            hkwargs["create_graph"] = torch.is_grad_enabled() and (_requires_grad(params) or molecule.rdm1.requires_grad)
        exc, xcfock, _ = xc_energy_and_grads(functional, params, molecule.rdm1, molecule, *args, **hkwargs)
        return energy + exc - (molecule.rdm1 * xcfock).sum() + coulomb_e + molecule.nuclear_repulsion

    return Harris_energy


molecule_predictor = energy_predictor  # name used in the notebooks' prose


# ---------------------------------------------------------------------------------------------------------
# losses and the training kernel  (grad_dft/train.py:312-359, 480-575; thin scalar glue over the predictor)
# ---------------------------------------------------------------------------------------------------------
def mse_energy_loss(params, compute_energy: Callable, atoms_list, truth_energies, elec_num_norm: bool = True, ranks=None) -> Array:
    """grad_dft/train.py:480-535: mean over molecules of ((E_pred - E_true) / n_elec)^2.  `compute_energy` is a
    non-SCF or SCF predictor returning a Molecule with `.energy`.  With `ranks=(rank, world)` only the molecules
    assigned to this rank (`distributed.shard_molecules`) are evaluated and the caller all-reduces loss and gradients
    (`distributed.allreduce_gradients`): the reference loops serially over the batch (train.py:519)."""
    if isinstance(atoms_list, Molecule):
        atoms_list = [atoms_list]
    idx = range(len(atoms_list))
    if ranks is not None:
        from .distributed import shard_molecules
        idx = shard_molecules([m.grid_size * m.ao.shape[1] ** 2 for m in atoms_list], *ranks)
    total = 0.0
    energy_only = getattr(compute_energy, "energy_only", None)
    batch = getattr(compute_energy, "energy_only_batch", None)
    idx = list(idx)
    batched = dict(zip(idx, batch(params, [atoms_list[i] for i in idx]))) if (batch is not None and len(idx) > 1) else {}
    for i in idx:
        atoms = atoms_list[i]
        if i in batched:
            energy = batched[i]
        else:
            energy = energy_only(params, atoms) if energy_only is not None else compute_energy(params, atoms).energy
        diff = energy - truth_energies[i]
        if elec_num_norm:
            num_elec = atoms.mo_occ.sum() if atoms.atom_index is None else (torch.as_tensor(atoms.atom_index).sum() - atoms.charge)
            diff = diff / num_elec
        total = total + diff ** 2
    return total / len(atoms_list)


def simple_energy_loss(params, compute_energy: Callable, atoms: Molecule, truth_energy):
    """grad_dft/train.py:537-560 (value_and_grad, has_aux): ((loss, E_pred), grads w.r.t. params)."""
    leaves = [p for p in _leaves(params) if p.requires_grad]
    energy_only = getattr(compute_energy, "energy_only", None)
    energy = energy_only(params, atoms) if energy_only is not None else compute_energy(params, atoms).energy
    loss = (energy - truth_energy) ** 2
    grads = torch.autograd.grad(loss, leaves)
    return (loss.detach(), energy.detach()), grads


def _leaves(params):
    if isinstance(params, torch.Tensor):
        return [params]
    if isinstance(params, dict):
        return [l for v in params.values() for l in _leaves(v)]
    if isinstance(params, (list, tuple)):
        return [l for v in params for l in _leaves(v)]
    return []


def train_kernel(tx: "torch.optim.Optimizer", loss: Callable) -> Callable:
    """grad_dft/train.py:312-359 with a torch optimizer standing in for the optax GradientTransformation:
    kernel(params, atoms, truth) -> (params, loss, predicted_energy); the optimizer holds the state."""

    def kernel(params, atoms, ground_truth_energy, *args):
        (cost_value, predicted), grads = loss(params, atoms, ground_truth_energy)
        leaves = [p for p in _leaves(params) if p.requires_grad]
        for p, g in zip(leaves, grads):
            p.grad = g
        tx.step()
        tx.zero_grad(set_to_none=True)
        return params, cost_value, predicted

    return kernel
