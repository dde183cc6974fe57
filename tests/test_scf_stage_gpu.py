"""The fused n x n tail of the DIIS SCF iteration (gdft_scf_diis_step / gdft_scf_occupy) against the step-by-step
restatement of grad_dft/evaluate.py:1111-1205 (JittableDiis), eigenproblem.py:125-129 and molecule.py:815-889 that the
larger molecules still run through the host framework -- cycle by cycle, through the out-of-bounds cycle (== max_diis) and
the shifting regime beyond it -- and the whole loop against the un-fused loop and the oracle."""
import pytest
import torch

import oracle
import graddft_b200 as gd
from graddft_b200 import evaluate, ops
from graddft_b200.molecule import get_occ, make_rdm1
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu
F64 = torch.float64


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


@pytest.mark.parametrize("n,m", [(12, 4), (43, 10), (64, 10), (7, 3)])
def test_diis_step_matches_the_host_framework_restatement(cuda_device, n, m):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(n)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64, device=dev)  # noqa: E731
    X = rn(n, n)
    S = X @ X.T / n + torch.eye(n, dtype=F64, device=dev)
    L_inv = evaluate.overlap_factor(S)
    diis = evaluate.JittableDiis(S, torch.eye(n, dtype=F64, device=dev), max_diis=m, A_is_identity=True)
    z = torch.zeros((m, 2, n, n), dtype=F64, device=dev)
    data = (z, z.clone(), torch.zeros(m, dtype=F64, device=dev), z.clone())
    fock_vec, err_vec, gram = torch.zeros_like(z), torch.zeros_like(z), torch.zeros((2, m, m), dtype=F64, device=dev)
    F0, D0 = rn(2, n, n), rn(2, n, n)
    F0, D0 = F0 + F0.transpose(1, 2), D0 + D0.transpose(1, 2)
    for cycle in range(2 * m + 4):
        # a slowly converging sequence, so that the CDIIS matrix stays well conditioned enough to compare solutions
        F = F0 + 0.3 ** min(cycle, 6) * (rn(2, n, n) + 0.0)
        F = 0.5 * (F + F.transpose(1, 2))
        D = D0 + 0.3 ** min(cycle, 6) * rn(2, n, n)
        D = 0.5 * (D + D.transpose(1, 2))
        F_ref, data = diis.run((D, F, torch.tensor(0.0, dtype=F64, device=dev)), data, cycle)
        C_ref = L_inv @ F_ref @ L_inv.T
        C, F_new, x = ops.scf_diis_step(cycle, F, D, S, L_inv, fock_vec, err_vec, gram)
        torch.cuda.synchronize()
        assert rel(F_new, F_ref) < 1e-8, (cycle, rel(F_new, F_ref))  # the 11 x 11 solve is the ill-conditioned part
        assert rel(C, L_inv @ F_new @ L_inv.T) < 1e-13, cycle
        assert abs(float(x.sum(1).max()) - 1.0) < 1e-8 and rel(C, C_ref) < 1e-8
        # ring buffers: same content as the reference's, in rotated physical order
        head = (cycle - m) % m if cycle > m else 0
        order = [(head + i) % m for i in range(m)]
        assert rel(err_vec[order], data[3]) < 1e-12 and torch.equal(fock_vec[order], data[1])
        G_ref = torch.einsum("iskl,jskl->sij", data[3], data[3])
        assert rel(gram[:, order][:, :, order], G_ref) < 1e-12


@pytest.mark.parametrize("n", [5, 43, 64])
def test_occupy_matches_the_host_framework_restatement(cuda_device, n):
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(100 + n)
    X = torch.randn(n, n, generator=g, dtype=F64, device=dev)
    L_inv = evaluate.overlap_factor(X @ X.T / n + torch.eye(n, dtype=F64, device=dev))
    A = torch.randn(2, n, n, generator=g, dtype=F64, device=dev)
    evals, V = torch.linalg.eigh(A + A.transpose(1, 2))
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(n)).to(dev)   # unsorted eigenvalues: ranks must still be right
    evals, V = evals[:, perm].contiguous(), V[:, :, perm].contiguous()
    evals[1, 1] = evals[1, 0]                                                        # an exact tie: stable order decides
    occ_prev = torch.zeros(2, n, dtype=F64, device=dev)
    occ_prev[0, : max(1, n // 3)] = 1.0
    occ_prev[1, : max(1, n // 4)] = 1.0
    mo_coeff, mo_occ, rdm1 = ops.scf_occupy(evals, V, L_inv, occ_prev)
    C_ref = L_inv.T @ V
    occ_ref = get_occ(evals, occ_prev.sum(1).round().to(torch.int64), n)
    assert rel(mo_coeff, C_ref) < 1e-13 and torch.equal(mo_occ, occ_ref)
    assert rel(rdm1, make_rdm1(C_ref, occ_ref)) < 1e-13


@pytest.mark.parametrize("cycles", [4, 14, 25])
def test_fused_loop_matches_the_unfused_loop_and_the_oracle(cuda_device, cycles, monkeypatch):
    mol = synthetic_molecule(3000, 21, n_omega=1, seed=1984, mask_frac=0.0)
    m = gd.molecule_from_tensors(mol, cuda_device)
    with torch.no_grad():
        monkeypatch.setenv("GDFT_SCF_FUSED", "1")
        a = gd.diff_scf_loop(gd.B3LYP, cycles=cycles)(None, m)
        monkeypatch.setenv("GDFT_SCF_FUSED", "0")
        b = gd.diff_scf_loop(gd.B3LYP, cycles=cycles)(None, m)
    if not bool(torch.isfinite(b.energy)):
        # upstream behaviour once the error vectors vanish: the CDIIS matrix turns singular and inv() yields NaN (both paths)
        assert not bool(torch.isfinite(a.energy))
        return
    assert abs(float(a.energy) - float(b.energy)) < 1e-8
    assert rel(a.fock, b.fock) < 1e-7 and rel(a.rdm1, b.rdm1) < 1e-6
    assert abs(float(a._norm_gorb) - float(b._norm_gorb)) < 1e-6 * max(1.0, float(b._norm_gorb))
    if cycles <= 14:
        e_ref, _ = oracle.diff_scf_loop_energy(mol, oracle.predict_b3lyp, cycles)
        assert abs(float(a.energy) - float(e_ref)) < 1e-7


def test_fused_loop_in_a_cuda_graph(cuda_device, monkeypatch):
    monkeypatch.setenv("GDFT_SCF_FUSED", "1")
    mol = synthetic_molecule(2000, 18, n_omega=1, seed=7, mask_frac=0.0)
    m = gd.molecule_from_tensors(mol, cuda_device)
    loop = gd.make_jitted_scf_loop(gd.B3LYP, cycles=12)
    with torch.no_grad():
        e_eager = float(gd.diff_scf_loop(gd.B3LYP, cycles=12)(None, m).energy)
        out = loop(None, m)
        out = loop(None, m)
    assert loop.last_call_was_graph
    assert abs(float(out.energy) - e_eager) < 1e-10


@pytest.mark.parametrize("n", [3, 43, 264, 700])
def test_aufbau_occupations_match_the_argsort_restatement(cuda_device, n):
    g = torch.Generator().manual_seed(n)
    ev = torch.randn(2, n, generator=g, dtype=F64)
    if n > 4:
        ev[0, 3] = ev[0, 1]  # exact ties: the stable order decides
        ev[1, n - 1] = ev[1, 0]
    occ_prev = torch.zeros(2, n, dtype=F64)
    occ_prev[0, : max(1, n // 3)] = 1.0
    occ_prev[1, : max(1, n // 5)] = 1.0
    want = get_occ(ev, occ_prev.sum(1).round().to(torch.int64), n)
    got = ops.aufbau_occupations(ev.to(cuda_device), occ_prev.to(cuda_device))
    assert torch.equal(got.cpu(), want)
    m = gd.molecule_from_tensors(synthetic_molecule(64, 12, seed=1), cuda_device)
    assert torch.equal(m.get_occ().cpu(), get_occ(m.mo_energy.cpu(), m.mo_occ.cpu().sum(1).round().to(torch.int64), 12))
