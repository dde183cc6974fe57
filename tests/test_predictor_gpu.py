"""energy_predictor / Functional.energy through the public API vs the CPU oracle (train.py:124-216 restated)."""
import pytest
import torch

import oracle
import graddft_b200 as gd
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu
F64 = torch.float64
E_TOL = 1e-8      # Ha, BASELINE.json / tests/integration/molecules/test_non_xc_energy.py:42
F_RTOL = 1e-7     # relative, BASELINE.json


def relerr(a, b):
    return float((a.cpu() - b).abs().max() / (b.abs().max() + 1e-300))


FUNCS = {"LSDA": gd.LSDA, "B88": gd.B88, "VWN": gd.VWN, "LYP": gd.LYP, "PW92": gd.PW92}


@pytest.mark.parametrize("name", list(FUNCS))
@pytest.mark.parametrize("N,n,seed", [(3000, 12, 1984), (2500, 43, 1993)])
def test_semilocal_predictor(cuda_device, name, N, n, seed):
    mol = synthetic_molecule(N, n, seed=seed, mask_frac=0.0)
    e_ref, f_ref = oracle.predict_semilocal(mol, name)
    m = gd.molecule_from_tensors(mol, cuda_device)
    e, f = gd.energy_predictor(FUNCS[name])(None, m)
    assert abs(float(e) - float(e_ref)) < E_TOL
    assert relerr(f, f_ref) < F_RTOL
    # Functional.energy + autograd wrt rdm1 (grad_dft/functional.py:255-288; notebook 02 cells 30-35)
    leaf = m.rdm1.clone().requires_grad_(True)
    e2 = FUNCS[name].energy(None, m.replace(rdm1=leaf))
    assert abs(float(e2) - float(e_ref)) < E_TOL
    (g,) = torch.autograd.grad(e2, leaf)
    D = mol["rdm1"].clone().requires_grad_(True)
    e_o = oracle.xc_energy_of_rdm1(D, mol, name) + oracle.nonXC(D.sum(0), mol["h1e"], mol["rep_tensor"], mol["nuclear_repulsion"])
    (g_ref,) = torch.autograd.grad(e_o, D)
    assert relerr(g, g_ref) < F_RTOL


@pytest.mark.parametrize("N,n,seed", [(3000, 12, 1984), (2500, 43, 1993)])
def test_b3lyp_predictor(cuda_device, N, n, seed):
    mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
    e_ref, f_ref = oracle.predict_b3lyp(mol)
    m = gd.molecule_from_tensors(mol, cuda_device)
    e, f = gd.energy_predictor(gd.B3LYP)(None, m)
    assert abs(float(e) - float(e_ref)) < E_TOL
    assert relerr(f, f_ref) < F_RTOL
    assert torch.equal(f, f.transpose(1, 2)) or relerr(f, f_ref.transpose(1, 2)) < F_RTOL


@pytest.mark.parametrize("N,n,seed", [(2000, 12, 1984), (1500, 30, 1993)])
def test_dm21_predictor(cuda_device, N, n, seed):
    mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
    params = oracle.dm21_mlp_init(seed=seed)
    e_ref, f_ref = oracle.predict_dm21(mol, params)
    m = gd.molecule_from_tensors(mol, cuda_device)
    fun = gd.DM21()
    p = {k: v.to(cuda_device) for k, v in params.items()}
    e, f = gd.energy_predictor(fun)(p, m)
    assert abs(float(e) - float(e_ref)) < E_TOL
    assert relerr(f, f_ref) < F_RTOL
    # gradient with respect to the network parameters (training step, train.py:312-359)
    pl = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    D = mol["rdm1"]
    e_o = oracle.xc_energy_of_rdm1(D, mol, "DM21", params=pl)
    g_ref = torch.autograd.grad(e_o, list(pl.values()))
    pd = {k: v.to(cuda_device).requires_grad_(True) for k, v in params.items()}
    e_x = fun.energy_xc_only(pd, m)
    g = torch.autograd.grad(e_x, list(pd.values()))
    for a, b, k in zip(g, g_ref, pd):
        assert relerr(a, b) < 1e-6 or float(b.abs().max()) < 1e-14, k


def test_masked_rows_are_inert_and_nan_free(cuda_device):
    """Rows with ao == 0 (rho == 0 exactly): reverse-mode autodiff of the reference formulas yields NaN there
    (0 * inf); the kernels return the finite forward-mode derivative, so such rows contribute exactly nothing."""
    N, n = 3000, 20
    mol = synthetic_molecule(N, n, n_omega=2, seed=1984, mask_frac=0.02)
    keep = mol["ao"].abs().sum(dim=1) > 0
    assert int((~keep).sum()) > 10
    sub = dict(mol)
    for k in ("ao", "grad_ao", "grad_n_ao2", "chi", "weights", "coords"):
        sub[k] = mol[k][keep].contiguous()
    e_ref, f_ref = oracle.predict_b3lyp(sub)
    m = gd.molecule_from_tensors(mol, cuda_device)
    e, f = gd.energy_predictor(gd.B3LYP)(None, m)
    assert bool(torch.isfinite(f).all())
    assert abs(float(e) - float(e_ref)) < E_TOL
    assert relerr(f, f_ref) < F_RTOL


def test_free_functions_and_errors(cuda_device):
    mol = synthetic_molecule(700, 9, n_omega=2, seed=1993)
    d = {k: (v.to(cuda_device) if isinstance(v, torch.Tensor) else v) for k, v in mol.items()}
    D = mol["rdm1"]
    assert relerr(gd.density(d["rdm1"], d["ao"]), oracle.density(D, mol["ao"])) < 1e-11
    assert relerr(gd.grad_density(d["rdm1"], d["ao"], d["grad_ao"]), oracle.grad_density(D, mol["ao"], mol["grad_ao"])) < 1e-11
    assert relerr(gd.kinetic_density(d["rdm1"], d["grad_ao"]), oracle.kinetic_density(D, mol["grad_ao"])) < 1e-11
    assert relerr(gd.lapl_density(d["rdm1"], d["ao"], d["grad_ao"], d["grad_n_ao2"]),
                  oracle.lapl_density(D, mol["ao"], mol["grad_ao"], mol["grad_n_ao2"])) < 1e-11
    assert relerr(gd.HF_energy_density(d["rdm1"], d["ao"], d["chi"]), oracle.HF_energy_density(D, mol["ao"], mol["chi"])) < 1e-11
    P = D.sum(0)
    assert relerr(gd.coulomb_potential(d["rdm1"].sum(0), d["rep_tensor"]), oracle.coulomb_potential(P, mol["rep_tensor"])) < 1e-11
    e = gd.nonXC(d["rdm1"].sum(0), d["h1e"], d["rep_tensor"], d["nuclear_repulsion"])
    assert abs(float(e) - float(oracle.nonXC(P, mol["h1e"], mol["rep_tensor"], mol["nuclear_repulsion"]))) < 1e-9
    with pytest.raises(TypeError):
        gd.density(d["rdm1"].float(), d["ao"])
    with pytest.raises(TypeError):
        gd.density(d["rdm1"][0], d["ao"])
    m = gd.molecule_from_tensors(mol, cuda_device)
    with pytest.raises(ValueError):
        m.HF_energy_density([0.3])
    with pytest.raises(ValueError):
        m.replace(chi=None, omegas=None).HF_energy_density([0.0])
    occ = m.get_occ()
    assert torch.equal(occ.cpu(), oracle.get_occ(mol["mo_energy"], mol["mo_occ"].sum(1).round().long()))
    assert relerr(m.make_rdm1(), oracle.make_rdm1(mol["mo_coeff"], mol["mo_occ"])) < 1e-13


def test_training_step_energy_loss(cuda_device):
    """Non-SCF training on a small batch (tests/integration/molecules/test_training.py:131-162 pattern): the loss is
    finite and >= 0, parameter gradients match torch-CPU autograd through the oracle, and 5 Adam steps reduce it."""
    dev = cuda_device
    mols = [synthetic_molecule(900 + 100 * i, 6 + 2 * i, n_omega=2, seed=1984 + i, mask_frac=0.0) for i in range(3)]
    truths = torch.tensor([-1.0, -2.0, -1.5], dtype=F64)
    fun = gd.DM21(layer_widths=(16, 16))
    flat = oracle.dm21_mlp_init(width=16, n_layers=2, seed=7)
    # oracle loss + gradient
    pl = {k: v.clone().requires_grad_(True) for k, v in flat.items()}
    loss_ref = 0.0
    for m, t in zip(mols, truths):
        e = oracle.xc_energy_of_rdm1(m["rdm1"], m, "DM21", params=pl) + oracle.nonXC(m["rdm1"].sum(0), m["h1e"], m["rep_tensor"], m["nuclear_repulsion"])
        loss_ref = loss_ref + ((e - t) / m["mo_occ"].sum()) ** 2
    loss_ref = loss_ref / 3
    g_ref = torch.autograd.grad(loss_ref, list(pl.values()))
    # kernels
    params = {k: v.to(dev).requires_grad_(True) for k, v in flat.items()}
    ms = [gd.molecule_from_tensors(m, dev) for m in mols]
    predictor = gd.non_scf_predictor(fun)
    loss = gd.mse_energy_loss(params, predictor, ms, truths.to(dev))
    assert abs(float(loss) - float(loss_ref)) < 1e-9 * max(1.0, abs(float(loss_ref)))
    g = torch.autograd.grad(loss, list(params.values()))
    for a, b, k in zip(g, g_ref, params):
        assert bool(torch.isfinite(a).all())
        assert relerr(a, b) < 1e-6 or float(b.abs().max()) < 1e-13, k
    # the batched entry (one network pass for the whole batch / per group of molecules) == the per-molecule entry
    for max_points in (10 ** 9, 1500):
        eb = predictor.energy_only_batch(params, ms, max_points=max_points)
        gb = torch.autograd.grad(sum(e * (k + 1.0) for k, e in enumerate(eb)), list(params.values()))
        es = [predictor.energy_only(params, m) for m in ms]
        gs = torch.autograd.grad(sum(e * (k + 1.0) for k, e in enumerate(es)), list(params.values()))
        for a, b in zip(eb, es):
            assert abs(float(a) - float(b)) < 1e-11 * max(1.0, abs(float(b)))
        for a, b in zip(gb, gs):
            assert relerr(a, b.cpu()) < 1e-9 or float(b.abs().max()) < 1e-13
    opt = torch.optim.Adam(list(params.values()), lr=1e-2)
    history = []
    for _ in range(5):
        opt.zero_grad()
        l = gd.mse_energy_loss(params, predictor, ms, truths.to(dev))
        l.backward()
        opt.step()
        history.append(float(l))
    assert history[-1] < history[0] and all(h >= 0 and h == h for h in history)


def _gapped_molecule(N, n, seed):
    """Synthetic molecule with a well-separated orbital spectrum, so that the SCF map is smooth in the parameters
    (aufbau occupations do not switch under a finite-difference step)."""
    mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
    mol["h1e"] = torch.diag(torch.linspace(-8.0, 8.0, n, dtype=F64)) + 0.05 * mol["h1e"]
    mol["rep_tensor"] = 0.05 * mol["rep_tensor"]
    mol["s1e"] = torch.eye(n, dtype=F64) + 0.2 * (mol["s1e"] - torch.eye(n, dtype=F64))
    return mol


def test_gradient_through_scf_loop_matches_finite_differences(cuda_device):
    """Training through the SCF loop (grad_dft/evaluate.py:917-1038 under jax.grad; examples/advanced_scripts/
    train_scf_loop.py): d E_scf / d params of a small semilocal neural functional (DM21 trunk on the 7 local inputs,
    LDA energy density) -- which differentiates every V_xc once more (second-order per-point kernels, the transposed
    density kernels, the safe eigh VJP, DIIS) -- against central finite differences along a random direction."""
    dev = cuda_device
    m = gd.molecule_from_tensors(_gapped_molecule(1200, 8, 1984), dev)
    fun = gd.DM21(layer_widths=(8, 8), nograd_densities=None, densitygrads=None, combine_densities=None, nograd_coefficient_inputs=None,
                  coefficient_input_grads=None, combine_inputs=None, local_features=1, needs_omegas=None)
    flat = fun.generate_DM21_weights(n_input_features=7, seed=3)
    gen = torch.Generator().manual_seed(5)
    direction = {k: torch.randn(v.shape, generator=gen, dtype=F64).to(dev) for k, v in flat.items()}
    for make_loop in (lambda: gd.diff_scf_loop(fun, cycles=3), lambda: gd.diff_simple_scf_loop(fun, cycles=3)):
        loop = make_loop()
        params = {k: v.to(dev).requires_grad_(True) for k, v in flat.items()}
        e = loop(params, m).energy
        grads = torch.autograd.grad(e, list(params.values()))
        slope = sum(float((g * direction[k]).sum()) for g, k in zip(grads, params))
        def central(h):
            with torch.no_grad():
                ep = loop({k: v.to(dev) + h * direction[k] for k, v in flat.items()}, m).energy
                em = loop({k: v.to(dev) - h * direction[k] for k, v in flat.items()}, m).energy
            return float(ep - em) / (2 * h)

        fd = (4.0 * central(1e-6) - central(2e-6)) / 3.0  # Richardson: the DIIS map is strongly curved in the parameters
        assert abs(slope - fd) < 5e-6 * max(1.0, abs(fd)), (slope, fd)
        assert abs(slope) > 1e-3  # the check is not vacuous


def test_gradient_through_scf_loop_hybrid_matches_traced_oracle(cuda_device):
    """Same for the hybrid DM21 functional, where jax.grad is NOT the total derivative (the HF energy density enters
    under stop_gradient, functional.py:176,203): the parameter gradient through 2 SCF cycles must equal torch-CPU
    autograd through the oracle's traced restatement of predict + loop, stop_gradients included."""
    dev = cuda_device
    mol = _gapped_molecule(700, 6, 1993)
    flat = oracle.dm21_mlp_init(width=8, n_layers=2, seed=3)
    fun = gd.DM21(layer_widths=(8, 8))
    m = gd.molecule_from_tensors(mol, dev)
    cases = ((lambda: gd.diff_simple_scf_loop(fun, cycles=2), oracle.diff_simple_scf_loop_energy),
             (lambda: gd.diff_scf_loop(fun, cycles=2), oracle.diff_scf_loop_energy))
    for make_loop, oracle_loop in cases:
        pl = {k: v.clone().requires_grad_(True) for k, v in flat.items()}
        e_ref, _ = oracle_loop(mol, lambda mm: oracle.predict_dm21_traced(mm, pl), 2)
        g_ref = torch.autograd.grad(e_ref, list(pl.values()))
        params = {k: v.to(dev).requires_grad_(True) for k, v in flat.items()}
        e = make_loop()(params, m).energy
        assert abs(float(e) - float(e_ref)) < 1e-7
        g = torch.autograd.grad(e, list(params.values()))
        scale = max(float(b.abs().max()) for b in g_ref)
        for a, b, k in zip(g, g_ref, params):
            assert float((a.cpu() - b).abs().max()) < 1e-6 * scale, k


def test_harris_energy(cuda_device):
    """grad_dft/train.py:220-308 against the same expression on the CPU oracle."""
    mol = synthetic_molecule(2000, 10, seed=1984, mask_frac=0.0)
    m = gd.molecule_from_tensors(mol, cuda_device)
    for name in ("LSDA", "B88", "LYP"):
        D = mol["rdm1"].clone().requires_grad_(True)
        exc = oracle.xc_energy_of_rdm1(D, mol, name)
        (v,) = torch.autograd.grad(exc, D)
        P = mol["rdm1"].sum(0)
        ref = ((mol["mo_occ"] * mol["mo_energy"]).sum() - oracle.coulomb_energy(P, mol["rep_tensor"]) + exc.detach()
               - (mol["rdm1"] * v).sum() + mol["nuclear_repulsion"])
        e = gd.Harris_energy_predictor(FUNCS[name])(None, m)
        assert abs(float(e) - float(ref)) < E_TOL, name


def test_harris_energy_parameter_gradient_matches_finite_differences(cuda_device):
    """jax.grad of the Harris energy (grad_dft/train.py:220-308) differentiates through V_xc: the -<rdm1, dV_xc/dtheta>
    term must be in the parameter gradient.  Semilocal neural functional and the hybrid DM21 (whose first-order tap path
    would return a constant V_xc), against central finite differences along a random direction."""
    dev = cuda_device
    m = gd.molecule_from_tensors(_gapped_molecule(1200, 8, 1984), dev)
    semilocal = gd.DM21(layer_widths=(8, 8), nograd_densities=None, densitygrads=None, combine_densities=None,
                        nograd_coefficient_inputs=None, coefficient_input_grads=None, combine_inputs=None, local_features=1, needs_omegas=None)
    hybrid = gd.DM21(layer_widths=(8, 8))
    for fun, nin in ((semilocal, 7), (hybrid, 11)):
        flat = fun.generate_DM21_weights(n_input_features=nin, seed=3)
        gen = torch.Generator().manual_seed(5)
        direction = {k: torch.randn(v.shape, generator=gen, dtype=F64).to(dev) for k, v in flat.items()}
        harris = gd.Harris_energy_predictor(fun)
        params = {k: v.to(dev).requires_grad_(True) for k, v in flat.items()}
        e = harris(params, m)
        grads = torch.autograd.grad(e, list(params.values()))
        slope = sum(float((g * direction[k]).sum()) for g, k in zip(grads, params))

        def central(h):
            with torch.no_grad():
                ep = harris({k: v.to(dev) + h * direction[k] for k, v in flat.items()}, m)
                em = harris({k: v.to(dev) - h * direction[k] for k, v in flat.items()}, m)
            return float(ep - em) / (2 * h)

        fd = (4.0 * central(1e-5) - central(2e-5)) / 3.0
        assert abs(slope - fd) < 1e-6 * max(1.0, abs(fd)), (slope, fd)
        # the term that used to be dropped is not negligible here: the E_xc-only slope differs from the full one
        exc = fun.energy_xc_only(params, m)
        slope_exc = sum(float((g * direction[k]).sum()) for g, k in zip(torch.autograd.grad(exc, list(params.values())), params))
        assert abs(slope - slope_exc) > 1e-6 * max(1.0, abs(fd))


@pytest.mark.parametrize("name", ["LSDA", "B88", "VWN", "LYP", "PW92", "B3LYP"])
def test_fused_xc_build_matches_the_generic_chain(cuda_device, name, monkeypatch):
    """The one-pass first-order XC build of the closed-form functionals (gdft_xc_point_fused behind train.xc_energy_and_grads)
    against the generic features -> combine -> clip -> quadrature -> autograd chain it replaces: E_xc, V_xc, the cotangent at the
    exact-exchange boundary and the whole predictor; masked (exactly zero density) rows included."""
    from graddft_b200.train import xc_energy_and_grads

    fun = dict(FUNCS, B3LYP=gd.B3LYP)[name]
    mol = synthetic_molecule(3001, 23, n_omega=1, seed=1993, mask_frac=0.01)
    m = gd.molecule_from_tensors(mol, cuda_device)
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("GDFT_FUSED_XC", mode)
        with torch.no_grad():
            exc, vxc, at = xc_energy_and_grads(fun, None, m.rdm1, m)
            e, f = gd.energy_predictor(fun)(None, m)
        res[mode] = (exc, vxc, at._memo()["xc_build"].g_densities if name == "B3LYP" else None, e, f)
    a, b = res["1"], res["0"]
    assert abs(float(a[0]) - float(b[0])) <= 1e-12 * abs(float(b[0]))
    assert relerr(a[1], b[1].cpu()) < 1e-12
    if name == "B3LYP":
        assert relerr(a[2], b[2].cpu()) < 1e-13
    assert abs(float(a[3]) - float(b[3])) < 1e-10 and relerr(a[4], b[4].cpu()) < 1e-11
    # against the oracle on a grid without exactly-zero densities (there the reference's own derivative is NaN: DESIGN.md section 4)
    monkeypatch.setenv("GDFT_FUSED_XC", "1")
    mol = synthetic_molecule(3001, 23, n_omega=1, seed=1994, mask_frac=0.0)
    m = gd.molecule_from_tensors(mol, cuda_device)
    with torch.no_grad():
        e, f = gd.energy_predictor(fun)(None, m)
    e_ref, f_ref = oracle.predict_b3lyp(mol) if name == "B3LYP" else oracle.predict_semilocal(mol, name)
    assert abs(float(e) - float(e_ref)) < E_TOL and relerr(f, f_ref) < F_RTOL
