"""Pins the CPU oracle (oracle/reference_math.py) to the golden vectors under tests/golden/*.npz, which were
produced by executing the reference's own source files (tests/golden/make_golden.py + jaxshim.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle

G = Path(__file__).resolve().parent / "golden"
F64 = torch.float64


def load(name):
    z = np.load(G / name)
    return {k: torch.from_numpy(z[k]) for k in z.files}


def close(a, b, rtol=1e-12, atol_scale=1e-13):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert torch.equal(torch.isfinite(a), torch.isfinite(b))  # incl. the reference's own NaNs (0 * inf at clipped points)
    fin = torch.isfinite(b)
    scale = float(b[fin].abs().max()) if bool(fin.any()) else 0.0
    err = (a[fin] - b[fin]).abs()
    assert bool((err <= rtol * b[fin].abs() + atol_scale * scale + 1e-300).all()), float(err.max())


@pytest.mark.parametrize("tag", ["a", "b"])
def test_molecule_ops(tag):
    d = load(f"molecule_ops_{tag}.npz")
    D, ao, gao, g2, chi = d["rdm1"], d["ao"], d["grad_ao"], d["grad_n_ao2"], d["chi"]
    close(oracle.density(D, ao), d["out_density"])
    close(oracle.grad_density(D, ao, gao), d["out_grad_density"])
    close(oracle.lapl_density(D, ao, gao, g2), d["out_lapl_density"])
    close(oracle.kinetic_density(D, gao), d["out_kinetic_density"])
    close(oracle.HF_energy_density(D, ao, chi), d["out_HF_energy_density"])
    P = D.sum(0)
    close(oracle.coulomb_potential(P, d["rep_tensor"]), d["out_coulomb_potential"])
    close(oracle.nonXC(P, d["h1e"], d["rep_tensor"], d["nuclear_repulsion"]), d["out_nonXC"])
    close(oracle.make_rdm1(d["mo_coeff"], d["mo_occ"]), d["out_make_rdm1"])
    assert torch.equal(oracle.get_occ(d["mo_energy"], d["mo_occ"].sum(1).round().long()), d["out_get_occ"])
    # VJP of the whole density family w.r.t. rdm1: torch autograd on the restatement AND the closed formula (a10)
    Dl = D.clone().requires_grad_(True)
    outs = [oracle.density(Dl, ao), oracle.grad_density(Dl, ao, gao), oracle.kinetic_density(Dl, gao),
            oracle.lapl_density(Dl, ao, gao, g2), oracle.HF_energy_density(Dl, ao, chi)]
    (g,) = torch.autograd.grad(sum((o * d[f"cot{i}"]).sum() for i, o in enumerate(outs)), Dl)
    close(g, d["out_density_family_vjp"], rtol=1e-11, atol_scale=1e-12)
    formula = oracle.density_vjp_formula(ao, gao, g2.sum(-1), d["cot0"], d["cot1"], d["cot2"], d["cot3"])
    formula = formula + oracle.HF_fock(chi, d["cot4"], ao).sum(0)
    close(formula, d["out_density_family_vjp"], rtol=1e-11, atol_scale=1e-12)


def test_pointwise():
    d = load("pointwise.npz")
    rho, grho, lapl, cot = d["rho"], d["grad_rho"], d["lapl"], d["cot"]
    cases = {
        "lsda_x_e": (lambda r, g, l: oracle.lsda_x_e(r), (0,)),
        "b88_x_e": (lambda r, g, l: oracle.b88_x_e(r, g), (0, 1)),
        "pw92_c_e": (lambda r, g, l: oracle.pw92_c_e(r), (0,)),
        "vwn_c_e": (lambda r, g, l: oracle.vwn_c_e(r), (0,)),
        "lyp_c_e": (lambda r, g, l: oracle.lyp_c_e(r, g, l), (0, 1, 2)),
    }
    names = ("rho", "grad_rho", "lapl")
    for name, (f, argn) in cases.items():
        leaves = [t.clone().requires_grad_(True) for t in (rho, grho, lapl)]
        out = f(*leaves)
        close(out.detach(), d[f"out_{name}"])
        grads = torch.autograd.grad((out * cot).sum(), [leaves[a] for a in argn])
        for a, g in zip(argn, grads):
            close(g, d[f"vjp_{name}_{names[a]}"], rtol=1e-10, atol_scale=1e-12)


def test_dm21_features():
    d = load("dm21_features.npz")
    D, ao, gao, chi = d["rdm1"], d["ao"], d["grad_ao"], d["chi"]
    rho, grho, tau = oracle.density(D, ao), oracle.grad_density(D, ao, gao), oracle.kinetic_density(D, gao)
    close(oracle.dm21_coefficient_inputs(rho, grho, tau), d["out_dm21_coefficient_inputs"])
    for t in ("LDA", "GGA", "MGGA"):
        close(oracle.dm21_densities(rho, grho, tau, t), d[f"out_dm21_densities_{t}"])
        close(oracle.mgga_feature_densities(rho, grho, tau, t), d[f"out_densities_{t}"])
    ehf = oracle.HF_energy_density(D, ao, chi)
    close(oracle.dm21_combine_cinputs(oracle.dm21_coefficient_inputs(rho, grho, tau), ehf), d["out_dm21_combine_cinputs"])
    close(oracle.dm21_combine_densities(oracle.dm21_densities(rho, grho, tau, "LDA"), ehf), d["out_dm21_combine_densities"])


@pytest.mark.parametrize("tag", ["a", "b"])
def test_predictor(tag):
    d = load(f"predictor_{tag}.npz")
    mol = {k: v for k, v in d.items() if not k.startswith(("energy_", "fock_", "functional_energy_", "densities_"))}
    for name in ("LSDA", "B88", "VWN", "LYP", "PW92"):
        e, f = oracle.predict_semilocal(mol, name)
        assert abs(float(e) - float(d[f"energy_{name}"])) < 1e-10, name
        assert abs(float(e) - float(d[f"functional_energy_{name}"])) < 1e-10, name
        close(f, d[f"fock_{name}"], rtol=1e-9, atol_scale=1e-12)
    e, f = oracle.predict_b3lyp(mol)
    assert abs(float(e) - float(d["energy_B3LYP"])) < 1e-10
    close(f, d["fock_B3LYP"], rtol=1e-9, atol_scale=1e-12)


def _wide_case(d, tag):
    """Regenerates the inputs of a predictor_wide.npz case (they are not stored) and checks them against the stored checksums."""
    from graddft_b200.synthetic import synthetic_molecule

    N, n, seed = (int(x) for x in d[f"{tag}_shape"])
    mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
    sums = torch.tensor([float(mol[k].double().sum()) for k in ("ao", "grad_ao", "rdm1", "weights", "rep_tensor", "chi", "h1e")], dtype=torch.float64)
    assert torch.allclose(sums, d[f"{tag}_checksums"], rtol=1e-12, atol=0), "synthetic_molecule no longer reproduces the golden inputs"
    return mol


def test_predictor_at_the_h2o_width():
    """The restatement against the reference's own energy_predictor at n = 43 (tests/golden/make_golden_wide.py)."""
    d = load("predictor_wide.npz")
    mol = _wide_case(d, "n43")
    for name in ("LSDA", "B88", "VWN", "LYP", "PW92"):
        e, f = oracle.predict_semilocal(mol, name)
        assert abs(float(e) - float(d[f"n43_energy_{name}"])) < 1e-9, name
        close(f, d[f"n43_fock_{name}"], rtol=1e-9, atol_scale=1e-12)
    e, f = oracle.predict_b3lyp(mol)
    assert abs(float(e) - float(d["n43_energy_B3LYP"])) < 1e-9
    close(f, d["n43_fock_B3LYP"], rtol=1e-9, atol_scale=1e-12)


def test_scf_loop_and_dm21_at_the_h2o_width():
    """diff_scf_loop (B3LYP: the hybrid route inside the DIIS loop; B88) and the DM21 predictor restated, against the reference's
    own evaluate.py / train.py at n = 43 (tests/golden/make_golden_wide.py -> scf_wide.npz)."""
    from graddft_b200.synthetic import synthetic_molecule

    d = load("scf_wide.npz")
    N, n, seed = (int(x) for x in d["shape"])
    mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
    sums = torch.tensor([float(mol[k].double().sum()) for k in ("ao", "grad_ao", "rdm1", "weights", "rep_tensor", "chi", "h1e")], dtype=torch.float64)
    assert torch.allclose(sums, d["checksums"], rtol=1e-12, atol=0)
    for name, cycles, pred in (("B3LYP", 3, oracle.predict_b3lyp), ("B88", 5, lambda m: oracle.predict_semilocal(m, "B88"))):
        e, out = oracle.diff_scf_loop_energy(mol, pred, cycles)
        assert abs(float(e) - float(d[f"diis_energy_{name}_{cycles}"])) < 1e-9, name
        close(out["rdm1"], d[f"diis_rdm1_{name}_{cycles}"], rtol=1e-7, atol_scale=1e-9)
    params = {k[len("param_"):]: v for k, v in d.items() if k.startswith("param_")}
    e, f = oracle.predict_dm21(mol, params)
    assert abs(float(e) - float(d["energy_DM21"])) < 1e-9
    close(f, d["fock_DM21"], rtol=1e-8, atol_scale=1e-11)


def _train_batch(d):
    from graddft_b200.synthetic import synthetic_molecule

    mols = []
    for (N, n, seed), z, sums in zip(d["shapes"].tolist(), d["atom_index"].tolist(), d["checksums"]):
        mol = synthetic_molecule(int(N), int(n), n_omega=2, seed=int(seed), mask_frac=0.0)
        got = torch.tensor([float(mol[k].double().sum()) for k in ("ao", "grad_ao", "rdm1", "weights", "rep_tensor", "chi", "h1e")], dtype=torch.float64)
        assert torch.allclose(got, sums, rtol=1e-12, atol=0), "synthetic_molecule no longer reproduces the golden inputs"
        mol["atom_index"] = torch.tensor([a for a in z if a > 0], dtype=torch.int64)
        mols.append(mol)
    return mols


def test_training_batch_loss_and_parameter_gradient():
    """mse_energy_loss over non_scf_predictor (train.py:480-535, evaluate.py:88-126) of the reference's own source for a
    three-molecule batch and a DM21-shaped network, with its gradient w.r.t. the parameters (train_batch.npz): the restatement
    (energy of the fixed density + torch autograd) reproduces both, with and without the electron-number normalisation."""
    d = load("train_batch.npz")
    mols = _train_batch(d)
    flat = {k[len("param_"):]: v for k, v in d.items() if k.startswith("param_")}
    for tag, norm in (("norm", True), ("plain", False)):
        pl = {k: v.clone().requires_grad_(True) for k, v in flat.items()}
        loss = 0.0
        for m, t in zip(mols, d["truths"]):
            e = oracle.xc_energy_of_rdm1(m["rdm1"], m, "DM21", params=pl) + oracle.nonXC(m["rdm1"].sum(0), m["h1e"], m["rep_tensor"], m["nuclear_repulsion"])
            diff = e - t
            if norm:
                diff = diff / m["atom_index"].sum()
            loss = loss + diff ** 2
        loss = loss / len(mols)
        assert abs(float(loss.detach()) - float(d[f"loss_{tag}"])) < 1e-10 * abs(float(d[f"loss_{tag}"]))
        for k, g in zip(pl, torch.autograd.grad(loss, list(pl.values()), allow_unused=True)):
            ref = d[f"grad_{tag}_{k}"]
            g = g if g is not None else torch.zeros_like(ref)
            assert float((g - ref).abs().max()) <= 1e-8 * float(ref.abs().max()) + 1e-14, (tag, k)


def _gapped(mol):
    """tests/golden/make_golden_wide.py::gapped."""
    n = mol["h1e"].shape[-1]
    mol["h1e"] = torch.diag(torch.linspace(-8.0, 8.0, n, dtype=torch.float64)) + 0.05 * mol["h1e"]
    mol["rep_tensor"] = 0.05 * mol["rep_tensor"]
    mol["s1e"] = torch.eye(n, dtype=torch.float64) + 0.2 * (mol["s1e"] - torch.eye(n, dtype=torch.float64))
    return mol


def test_parameter_gradient_through_the_scf_loops():
    """jax.grad through diff_scf_loop / diff_simple_scf_loop (evaluate.py:917-1038, 257-352) of the reference's own source for the
    hybrid DM21 functional (scf_grad.npz): the traced restatement -- stop_gradients, safe-eigh VJP and DIIS included -- gives the
    same energy and the same gradient w.r.t. every parameter."""
    from graddft_b200.synthetic import synthetic_molecule

    d = load("scf_grad.npz")
    N, n, seed = (int(x) for x in d["shape"])
    mol = _gapped(synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0))
    sums = torch.tensor([float(mol[k].double().sum()) for k in ("ao", "grad_ao", "rdm1", "weights", "rep_tensor", "chi", "h1e")], dtype=torch.float64)
    assert torch.allclose(sums, d["checksums"], rtol=1e-12, atol=0)
    flat = {k[len("param_"):]: v for k, v in d.items() if k.startswith("param_")}
    for tag, loop in (("diis", oracle.diff_scf_loop_energy), ("simple", oracle.diff_simple_scf_loop_energy)):
        pl = {k: v.clone().requires_grad_(True) for k, v in flat.items()}
        e, _ = loop(mol, lambda mm: oracle.predict_dm21_traced(mm, pl), 2)
        assert abs(float(e.detach()) - float(d[f"energy_{tag}"])) < 1e-9, tag
        grads = torch.autograd.grad(e, list(pl.values()), allow_unused=True)
        scale = max(float(d[f"grad_{tag}_{k}"].abs().max()) for k in pl)
        for k, g in zip(pl, grads):
            ref = d[f"grad_{tag}_{k}"]
            g = g if g is not None else torch.zeros_like(ref)
            assert float((g - ref).abs().max()) < 1e-7 * scale, (tag, k)


def test_predictor_dm21():
    d = load("predictor_dm21.npz")
    mol = {k: v for k, v in d.items() if not k.startswith(("energy_", "fock_", "param_", "out_"))}
    params = {k[len("param_"):]: v for k, v in d.items() if k.startswith("param_")}
    D, ao, gao, chi = mol["rdm1"], mol["ao"], mol["grad_ao"], mol["chi"]
    rho, grho, tau = oracle.density(D, ao), oracle.grad_density(D, ao, gao), oracle.kinetic_density(D, gao)
    ci = oracle.dm21_combine_cinputs(oracle.dm21_coefficient_inputs(rho, grho, tau), oracle.HF_energy_density(D, ao, chi))
    close(ci, d["out_cinputs"])
    close(oracle.dm21_mlp(params, ci), d["out_coefficients"], rtol=1e-10, atol_scale=1e-12)
    e, f = oracle.predict_dm21(mol, params)
    assert abs(float(e) - float(d["energy_DM21"])) < 1e-10
    close(f, d["fock_DM21"], rtol=1e-8, atol_scale=1e-11)


def test_scf_loops():
    """diff_scf_loop / JittableDiis / safe_fock_solver restatement vs the reference's own evaluate.py."""
    d = load("scf_loops.npz")
    mol = {k: v for k, v in d.items() if not k.startswith(("diis_", "simple_"))}
    for name, cycles in (("B88", 4), ("LSDA", 12)):
        e, out = oracle.diff_scf_loop_energy(mol, lambda m: oracle.predict_semilocal(m, name), cycles)
        assert abs(float(e) - float(d[f"diis_energy_{name}_{cycles}"])) < 1e-8, (name, float(e), float(d[f"diis_energy_{name}_{cycles}"]))
        close(out["rdm1"], d[f"diis_rdm1_{name}_{cycles}"], rtol=1e-6, atol_scale=1e-8)
