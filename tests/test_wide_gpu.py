"""Predictor- and SCF-level parity at the widths BASELINE.json names (n = 160 and the C3 width n = 264), where the tile
classes and K-split logic of the density kernels differ from the small-n cases of test_predictor_gpu.py: B3LYP
energy_predictor, a 2-cycle diff_scf_loop / make_jitted_scf_loop and the DM21 predictor through the public API against the
CPU oracle (grad_dft/train.py:124-216, grad_dft/evaluate.py:917-1038 restated).  The rep_tensor (38.9 GB dense at
n = 264) lives densely on the GPU only; the oracle contracts the same tensor in its factorised form.
Tolerances: BASELINE.json (|dE| < 1e-8 Ha, Fock 1e-7 relative) = tests/integration/molecules/test_non_xc_energy.py:42."""
import math

import pytest
import torch

import oracle
import graddft_b200 as gd
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu
F64 = torch.float64
E_TOL, F_RTOL = 1e-8, 1e-7


def relerr(a, b):
    return float((a.cpu() - b).abs().max() / (b.abs().max() + 1e-300))


def wide_molecule(N, n, seed, device):
    """(host dict for the oracle with a FactorizedERI, device Molecule with the dense tensor)."""
    mol = synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0, with_eri=False)
    g = torch.Generator().manual_seed(seed + 7)
    Q = 2 * n
    B = torch.randn(Q, n, n, generator=g, dtype=F64)
    B = 0.5 * (B + B.transpose(1, 2))
    eri = oracle.FactorizedERI(B, Q / 0.05)  # scaled like test_predictor_gpu._gapped_molecule: |E| stays O(100) Ha
    # a well-conditioned spectrum (occupations must not flip between two round-off-different evaluations)
    mol["h1e"] = torch.diag(torch.linspace(-6.0, 6.0, n, dtype=F64)) + 0.05 * mol["h1e"]
    host = dict(mol, rep_tensor=eri)
    dev = dict(mol)
    dev["rep_tensor"] = eri.dense(device)
    return host, gd.molecule_from_tensors(dev, device)


@pytest.mark.parametrize("N,n,seed", [(3000, 160, 1984), (3100, 264, 1993)])
def test_b3lyp_predictor_wide(cuda_device, N, n, seed):
    host, m = wide_molecule(N, n, seed, cuda_device)
    e_ref, f_ref = oracle.predict_b3lyp(host)
    e, f = gd.energy_predictor(gd.B3LYP)(None, m)
    assert abs(float(e) - float(e_ref)) < E_TOL, (float(e), float(e_ref))
    assert relerr(f, f_ref) < F_RTOL


@pytest.mark.parametrize("N,n,seed", [(3000, 160, 1984), (3100, 264, 1993)])
def test_scf_loop_wide(cuda_device, N, n, seed):
    """Two DIIS cycles (each: extrapolation, generalised eigenproblem at this width, occupations, rdm1, Fock build), eager
    and through make_jitted_scf_loop, against the oracle's loop."""
    host, m = wide_molecule(N, n, seed, cuda_device)
    e_ref, mol_ref = oracle.diff_scf_loop_energy(host, oracle.predict_b3lyp, 2)
    with torch.no_grad():
        out = gd.diff_scf_loop(gd.B3LYP, cycles=2)(None, m)
        assert abs(float(out.energy) - float(e_ref)) < E_TOL, (float(out.energy), float(e_ref))
        assert relerr(out.rdm1, mol_ref["rdm1"]) < 1e-6
        assert relerr(out.fock, mol_ref["fock"]) < F_RTOL
        jit = gd.make_jitted_scf_loop(gd.B3LYP, cycles=2)
        for _ in range(2):  # first call captures, second replays
            out2 = jit(None, m)
            assert abs(float(out2.energy) - float(e_ref)) < E_TOL
            assert relerr(out2.fock, mol_ref["fock"]) < F_RTOL


@pytest.mark.parametrize("N,n,seed", [(2600, 160, 1984), (2700, 264, 1993)])
def test_dm21_predictor_wide(cuda_device, N, n, seed):
    host, m = wide_molecule(N, n, seed, cuda_device)
    params = oracle.dm21_mlp_init(seed=seed)
    e_ref, f_ref = oracle.predict_dm21(host, params)
    p = {k: v.to(cuda_device) for k, v in params.items()}
    e, f = gd.energy_predictor(gd.DM21())(p, m)
    assert abs(float(e) - float(e_ref)) < E_TOL, (float(e), float(e_ref))
    assert relerr(f, f_ref) < F_RTOL
