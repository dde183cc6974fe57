"""Generates tests/golden/*.npz by executing the reference's own source files (imported from /root/reference,
unmodified) on the torch-backed jax stand-in of jaxshim.py.  Run here (the build container), commit the vectors:

    python tests/golden/make_golden.py

Each file holds the inputs and the reference outputs of one slice of the hot path: the grid/AO free functions,
the closed-form energy densities with their VJPs, the DM21 feature functions, and energy_predictor's (energy,
Fock) for LSDA/B88/VWN/LYP/PW92/B3LYP and for DM21 with seeded weights.  /root/reference does not exist on the
GPU box; nothing under tests/ reads it at test time.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import jaxshim  # noqa: E402
from graddft_b200.synthetic import synthetic_molecule  # noqa: E402

F64 = torch.float64
gd = jaxshim.install()
J = jaxshim._j


def np_(t):
    return t.detach().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)


def ref_molecule(mol, omegas=None):
    g2 = {2: J(mol["grad_n_ao2"])} if "grad_n_ao2" in mol else None
    return gd.Molecule(
        grid=gd.Grid(J(mol["coords"]), J(mol["weights"])), atom_index=J(torch.zeros(1, dtype=torch.int64)), nuclear_pos=J(torch.zeros(1, 3, dtype=F64)),
        ao=J(mol["ao"]), grad_ao=J(mol["grad_ao"]), grad_n_ao=g2, rdm1=J(mol["rdm1"]), nuclear_repulsion=J(mol["nuclear_repulsion"]),
        h1e=J(mol["h1e"]), vj=None, mo_coeff=J(mol["mo_coeff"]), mo_occ=J(mol["mo_occ"]), mo_energy=J(mol["mo_energy"]),
        s1e=J(mol["s1e"]), omegas=(J(mol["omegas"]) if "omegas" in mol else None), chi=(J(mol["chi"]) if "chi" in mol else None),
        rep_tensor=J(mol["rep_tensor"]),
    )


def grid_quantities(N, seed):
    g = torch.Generator().manual_seed(seed)
    rho = torch.exp(-14.0 * torch.rand(N, 2, generator=g, dtype=F64)) * 3.0
    grho = torch.randn(N, 2, 3, generator=g, dtype=F64) * rho[:, :, None] ** (4.0 / 3.0)
    tau = torch.rand(N, 2, generator=g, dtype=F64) * rho ** (5.0 / 3.0) * 3.0
    lapl = torch.randn(N, 2, generator=g, dtype=F64) * rho
    rho[:4, 0] = 1e-31   # one channel under the clip
    rho[4:8] = 1e-33     # both under the clip
    rho[8:12] = 3e-30    # just above
    return rho, grho, tau, lapl


def main():
    out = {}
    # ---- 1. grid / AO free functions -------------------------------------------------------------------
    for tag, (N, n, seed, sym) in {"a": (257, 7, 1984, True), "b": (190, 12, 1993, False)}.items():
        mol = synthetic_molecule(N, n, n_omega=2, seed=seed, symmetric_rdm1=sym, mask_frac=0.0)
        m = ref_molecule(mol)
        d = {k: np_(v) for k, v in mol.items()}
        d["out_density"] = np_(m.density())
        d["out_grad_density"] = np_(m.grad_density())
        d["out_lapl_density"] = np_(m.lapl_density())
        d["out_kinetic_density"] = np_(m.kinetic_density())
        d["out_HF_energy_density"] = np_(m.HF_energy_density(m.omegas))
        d["out_coulomb_potential"] = np_(m.get_coulomb_potential())
        d["out_nonXC"] = np_(m.nonXC())
        d["out_make_rdm1"] = np_(m.make_rdm1())
        d["out_get_occ"] = np_(m.get_occ())
        # reference VJP of the density family w.r.t. rdm1 (what value_and_grad differentiates through)
        gcot = torch.Generator().manual_seed(seed + 1)
        cots = [torch.randn(N, 2, generator=gcot, dtype=F64), torch.randn(N, 2, 3, generator=gcot, dtype=F64),
                torch.randn(N, 2, generator=gcot, dtype=F64), torch.randn(N, 2, generator=gcot, dtype=F64),
                torch.randn(2, 2, N, generator=gcot, dtype=F64)]

        def contracted(rdm1):
            mm = m.replace(rdm1=rdm1)
            outs = [mm.density(), mm.grad_density(), mm.kinetic_density(), mm.lapl_density(), mm.HF_energy_density(mm.omegas)]
            return sum((o * J(c)).sum() for o, c in zip(outs, cots))

        d["out_density_family_vjp"] = np_(jaxshim.grad(contracted)(m.rdm1))
        for i, c in enumerate(cots):
            d[f"cot{i}"] = np_(c)
        np.savez_compressed(HERE / f"molecule_ops_{tag}.npz", **d)

    # ---- 2. closed-form energy densities + VJPs ---------------------------------------------------------
    pf = sys.modules["grad_dft.popular_functionals"]
    fn = sys.modules["grad_dft.functional"]
    rho, grho, tau, lapl = grid_quantities(600, 1984)
    d = {"rho": np_(rho), "grad_rho": np_(grho), "tau": np_(tau), "lapl": np_(lapl)}
    cases = {
        "lsda_x_e": (lambda r, g, l: pf.lsda_x_e(r, 1e-30), (0,)),
        "b88_x_e": (lambda r, g, l: pf.b88_x_e(r, g), (0, 1)),
        "pw92_c_e": (lambda r, g, l: pf.pw92_c_e(r), (0,)),
        "vwn_c_e": (lambda r, g, l: pf.vwn_c_e(r), (0,)),
        "lyp_c_e": (lambda r, g, l: pf.lyp_c_e(r, g, l), (0, 1, 2)),
    }
    gcot = torch.Generator().manual_seed(77)
    cot = torch.randn(600, generator=gcot, dtype=F64)
    d["cot"] = np_(cot)
    for name, (f, argn) in cases.items():
        d[f"out_{name}"] = np_(f(J(rho), J(grho), J(lapl)))
        grads = jaxshim.grad(lambda r, g, l: (f(r, g, l) * J(cot)).sum(), argnums=argn)(J(rho), J(grho), J(lapl))
        for a, gr in zip(argn, grads):
            d[f"vjp_{name}_{('rho', 'grad_rho', 'lapl')[a]}"] = np_(gr)
    np.savez_compressed(HERE / "pointwise.npz", **d)

    # ---- 3. DM21 feature functions ----------------------------------------------------------------------
    mol = synthetic_molecule(211, 9, n_omega=2, seed=1993, mask_frac=0.0)
    m = ref_molecule(mol)
    d = {k: np_(v) for k, v in mol.items()}
    d["out_dm21_coefficient_inputs"] = np_(fn.dm21_coefficient_inputs(m))
    for t in ("LDA", "GGA", "MGGA"):
        d[f"out_dm21_densities_{t}"] = np_(fn.dm21_densities(m, functional_type=t))
    ehf = m.HF_energy_density(m.omegas)
    d["out_dm21_combine_cinputs"] = np_(fn.dm21_combine_cinputs(fn.dm21_coefficient_inputs(m), ehf))
    d["out_dm21_combine_densities"] = np_(fn.dm21_combine_densities(fn.dm21_densities(m), ehf))
    for t in ("LDA", "GGA", "MGGA"):
        d[f"out_densities_{t}"] = np_(fn.densities(m, functional_type=t))
    np.savez_compressed(HERE / "dm21_features.npz", **d)

    # ---- 4. energy_predictor for the closed-form functionals --------------------------------------------
    for tag, (N, n, seed) in {"a": (300, 7, 1984), "b": (220, 12, 1993)}.items():
        mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
        m = ref_molecule(mol)
        d = {k: np_(v) for k, v in mol.items()}
        for name in ("LSDA", "B88", "VWN", "LYP", "PW92", "B3LYP"):
            functional = getattr(gd, name)
            e, fock = gd.energy_predictor(functional)(None, m)
            d[f"energy_{name}"], d[f"fock_{name}"] = np_(e), np_(fock)
            d[f"functional_energy_{name}"] = np_(functional.energy(None, m))
            d[f"densities_{name}"] = np_(functional.compute_densities(m))
        np.savez_compressed(HERE / f"predictor_{tag}.npz", **d)

    # ---- 5. DM21 predictor with seeded weights -----------------------------------------------------------
    import oracle
    mol = synthetic_molecule(180, 8, n_omega=2, seed=1984, mask_frac=0.0)
    m = ref_molecule(mol)
    flat = oracle.dm21_mlp_init(width=32, n_layers=3, seed=1984)
    tree = {}
    for k, v in flat.items():
        layer, leaf = k.split(".")
        tree.setdefault(layer, {})[leaf] = J(v)
    dm21 = gd.DM21()
    dm21.layer_widths = [32, 32, 32]
    e, fock = gd.energy_predictor(dm21)({"params": tree}, m)
    d = {k: np_(v) for k, v in mol.items()}
    for k, v in flat.items():
        d["param_" + k] = np_(v)
    d["energy_DM21"], d["fock_DM21"] = np_(e), np_(fock)
    cin = dm21.compute_coefficient_inputs(m)
    d["out_cinputs"] = np_(cin)
    d["out_coefficients"] = np_(dm21.apply({"params": tree}, cin))
    np.savez_compressed(HERE / "predictor_dm21.npz", **d)
    # ---- 6. jitted SCF loops (evaluate.py:257-352, 917-1038) ------------------------------------------------
    mol = synthetic_molecule(260, 8, n_omega=0, seed=1993, mask_frac=0.0, with_grad2=False)
    m = ref_molecule(mol)
    d = {k: np_(v) for k, v in mol.items()}
    for name, cycles in (("B88", 4), ("LSDA", 12)):
        out = gd.diff_scf_loop(getattr(gd, name), cycles=cycles)(None, m)
        d[f"diis_energy_{name}_{cycles}"], d[f"diis_rdm1_{name}_{cycles}"], d[f"diis_fock_{name}_{cycles}"] = np_(out.energy), np_(out.rdm1), np_(out.fock)
    out = gd.diff_simple_scf_loop(gd.LSDA, cycles=3, mixing_factor=0.4)(None, m)
    d["simple_energy_LSDA_3"], d["simple_rdm1_LSDA_3"] = np_(out.energy), np_(out.rdm1)
    np.savez_compressed(HERE / "scf_loops.npz", **d)
    for p in sorted(HERE.glob("*.npz")):
        print(p.name, p.stat().st_size)


if __name__ == "__main__":
    main()
