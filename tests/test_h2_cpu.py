"""The oracle's from-scratch H2/STO-3G builder (oracle/h2_sto3g.py) against published numbers:
Szabo & Ostlund, Modern Quantum Chemistry, section 3.5.2 (R = 1.4 bohr), and against grid identities evaluated with the
oracle's restatement of grad_dft/molecule.py.  The analogue of tests/integration/molecules/test_non_xc_energy.py:42,138-222
without PySCF."""
import numpy as np
import pytest
import torch

import oracle
from oracle import h2_sto3g


@pytest.fixture(scope="module")
def h2():
    return h2_sto3g.build_h2()


def test_integrals_match_szabo_ostlund(h2):
    e, so = h2["expected"], h2_sto3g.SZABO_OSTLUND
    V = e["V_per_nucleus"]
    got = {"S12": e["S"][0, 1], "T11": e["T"][0, 0], "T12": e["T"][0, 1], "V11_one_centre": V[0, 0, 0], "V12_one_centre": V[0, 0, 1],
           "V22_at_centre1": V[0, 1, 1], "1111": e["eri"][0, 0, 0, 0], "1122": e["eri"][0, 0, 1, 1], "2111": e["eri"][1, 0, 0, 0],
           "2121": e["eri"][1, 0, 1, 0], "E_RHF": e["E_RHF"]}
    for k, v in so.items():
        assert abs(got[k] - v) < 6e-5, (k, got[k], v)  # the book prints four decimals
    assert abs(e["S"][0, 0] - 1.0) < 1e-6  # STO-3G contraction is normalised to ~1e-7
    g = e["eri"]
    assert np.allclose(g, g.transpose(1, 0, 2, 3)) and np.allclose(g, g.transpose(2, 3, 0, 1))


def test_grid_identities_with_the_oracle(h2):
    e, w = h2["expected"], h2["weights"]
    D = h2["rdm1"]
    rho = oracle.density(D, h2["ao"])
    assert abs(float((w[:, None] * rho).sum()) - 2.0) < 1e-7
    tau = oracle.kinetic_density(D, h2["grad_ao"])
    assert abs(float((w[:, None] * tau).sum()) - e["kinetic"]) < 1e-7
    lap = oracle.lapl_density(D, h2["ao"], h2["grad_ao"], h2["grad_n_ao2"])
    assert abs(float((w[:, None] * lap).sum())) < 1e-6
    g = oracle.grad_density(D, h2["ao"], h2["grad_ao"])
    assert float((w[:, None, None] * g).sum(0).abs().max()) < 1e-7  # integral of a gradient of a bound density vanishes
    ehf = oracle.HF_energy_density(D, h2["ao"], h2["chi"])
    assert abs(float((ehf[0] * w).sum()) - e["E_x_HF"]) < 1e-7
    assert float((ehf[1] * w).sum()) > float((ehf[0] * w).sum())  # erf-attenuated exchange is weaker
    non_xc = float(oracle.nonXC(D.sum(0), h2["h1e"], h2["rep_tensor"], h2["nuclear_repulsion"]))
    assert abs(non_xc - e["nonXC"]) < 1e-12
    assert abs(non_xc + float((ehf[0] * w).sum()) - h2_sto3g.SZABO_OSTLUND["E_RHF"]) < 6e-5
    # the orbitals are S-orthonormal and the density matrix holds two electrons
    C, S = h2["mo_coeff"][0], h2["s1e"]
    assert torch.allclose(C.T @ S @ C, torch.eye(2, dtype=torch.float64), atol=1e-12)
    assert abs(float((D.sum(0) * S).sum()) - 2.0) < 1e-12


def test_grid_converges():
    coarse, fine = h2_sto3g.build_h2(n_rad=40, n_theta=20, n_phi=4), h2_sto3g.build_h2(n_rad=90, n_theta=40, n_phi=6)
    ex = []
    for m in (coarse, fine):
        rho = oracle.density(m["rdm1"], m["ao"])
        ex.append(float(oracle.integrate(oracle.lsda_x_e(rho).sum(1) if oracle.lsda_x_e(rho).dim() > 1 else oracle.lsda_x_e(rho), m["weights"])))
    assert abs(ex[0] - ex[1]) < 1e-5 and -0.7 < ex[1] < -0.5  # Dirac exchange of H2: about -0.57 Ha at this density
