"""graddft_b200 -- B200-native kernels behind Grad DFT's per-SCF-iteration hot path.

Public names follow grad_dft/__init__.py:15-84.  Importing the package does not load the CUDA library;
the first kernel call does, and raises if libgdft_b200.so is missing (there is no CPU fallback).
"""
from .molecule import (  # noqa: F401
    Grid, Molecule, Reaction, abs_clip, coulomb_energy, coulomb_potential, density, get_occ, grad_density, HF_energy_density,
    HF_density_grad_2_Fock, HF_coefficient_input_grad_2_Fock, kinetic_density, lapl_density, make_rdm1, molecule_from_tensors, nonXC,
    one_body_energy, orbital_grad,
)
from .functional import (  # noqa: F401
    DM21, Functional, NeuralFunctional, canonicalize_inputs, dm21_coefficient_inputs, dm21_combine_cinputs, dm21_combine_densities,
    dm21_densities, dm21_hfgrads_cinputs, dm21_hfgrads_densities, densities, stop_gradient,
)
from .popular_functionals import B3LYP, B88, LSDA, LYP, PW92, VWN  # noqa: F401
from .train import Harris_energy_predictor, energy_predictor, molecule_predictor, mse_energy_loss, simple_energy_loss, train_kernel, xc_energy_and_grads  # noqa: F401
from .evaluate import (  # noqa: F401
    JittableDiis, non_scf_predictor, diff_scf_loop, diff_simple_scf_loop, make_jitted_scf_loop, make_simple_scf_loop, safe_eigh, safe_fock_solver,
)
from .interface import Archive, generate_chi_tensor, loader, make_reaction, save_molecule_data, saver  # noqa: F401,E402
