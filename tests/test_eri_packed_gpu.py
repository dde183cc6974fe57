"""Packed rep_tensor (include/gdft_b200.h "packed rep_tensor"): the pair-symmetric quarter of the J sweep of
grad_dft/molecule.py:788-811 against the oracle and against the plain sweep; symmetry detection, sharded pair rows,
autograd closure and the packing policy."""
import os

import pytest
import torch

import oracle
from graddft_b200 import distributed as gdist
from graddft_b200 import ops
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu
F64 = torch.float64


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(autouse=True)
def _fresh_policy():
    ops._PACKED_ERI.clear()
    gdist._PACKED_BLOCKS.clear()
    old = os.environ.get("GDFT_PACK_ERI")
    yield
    if old is None:
        os.environ.pop("GDFT_PACK_ERI", None)
    else:
        os.environ["GDFT_PACK_ERI"] = old


@pytest.mark.parametrize("n", [17, 24, 43])
def test_packed_j_matches_oracle_and_plain_sweep(cuda_device, n):
    mol = synthetic_molecule(64, n, seed=1984)
    eri, P = mol["rep_tensor"].to(cuda_device), mol["rdm1"].sum(0).to(cuda_device)
    P = P + 0.1 * torch.randn(n, n, dtype=F64, device=cuda_device)  # not symmetric: the packed sweep must symmetrise it itself
    asym, big = ops.eri_symmetry_defect(eri, n)
    assert asym == 0.0 and big > 0
    pe = ops.PackedERI.from_rows(eri, n)
    assert pe.complete and pe.pairs == n * (n + 1) // 2
    J, EJ = pe.coulomb(P, want_energy=True)
    J_ref = oracle.coulomb_potential(P.cpu(), mol["rep_tensor"])
    assert rel(J.cpu(), J_ref) < 1e-13
    assert abs(float(EJ) - float(oracle.coulomb_energy(P.cpu(), mol["rep_tensor"]))) < 1e-11 * abs(float(EJ))
    assert rel(J, ops.coulomb_j(P, eri)) < 1e-13
    assert torch.equal(J, J.T)  # both triangles come from the same packed entry
    assert torch.equal(pe.coulomb(P), J)  # run-to-run reproducible


def test_symmetry_detection_and_policy(cuda_device):
    n = 20
    mol = synthetic_molecule(32, n, seed=3)
    eri = mol["rep_tensor"].to(cuda_device)
    P = mol["rdm1"].sum(0).to(cuda_device)
    os.environ["GDFT_PACK_ERI"] = "auto"
    assert ops.packed_eri_for(eri) is None           # first use: the plain sweep
    pe = ops.packed_eri_for(eri)                     # second use: packed
    assert pe is not None and pe.exchange_symmetric
    assert ops.packed_eri_for(eri) is pe
    eri[3, 5, 7, 2] += 1e-6                          # in-place edit: version counter changes, the packed copy is not reused
    os.environ["GDFT_PACK_ERI"] = "always"
    assert ops.packed_eri_for(eri) is None           # ... and the edited tensor is no longer symmetric: refused
    asym, big = ops.eri_symmetry_defect(eri, n)
    assert abs(asym - 1e-6) < 1e-12
    bad = mol["rep_tensor"].to(cuda_device).clone()
    bad[:] = bad + 1e-9 * torch.randn_like(bad)
    assert ops.packed_eri_for(bad) is None
    J = ops.coulomb_j_auto(P, bad)                   # falls back to the sweep of the tensor as given
    assert rel(J.cpu(), oracle.coulomb_potential(P.cpu(), bad.cpu())) < 1e-13
    os.environ["GDFT_PACK_ERI"] = "never"
    assert ops.packed_eri_for(mol["rep_tensor"].to(cuda_device)) is None
    assert ops.packed_eri_for(synthetic_molecule(8, 8, seed=1)["rep_tensor"].to(cuda_device)) is None  # below PACK_ERI_MIN_N


def test_row_blocks_and_pair_rows_assemble_the_full_j(cuda_device):
    n, world = 29, 3
    mol = synthetic_molecule(32, n, seed=11)
    eri, P = mol["rep_tensor"].to(cuda_device), mol["rdm1"].sum(0).to(cuda_device)
    full = ops.PackedERI.from_rows(eri, n).coulomb(P)
    flat = eri.reshape(n * n, n, n)
    # contiguous (p,q) row blocks: each block packs the pair rows it contains
    acc, seen = torch.zeros_like(full), 0
    for r in range(world):
        r0, r1 = gdist.shard_bounds(n * n, r, world, align=32)
        pe = ops.PackedERI.from_rows(flat[r0:r1].contiguous(), n, r0)
        seen += pe.pairs
        acc += pe.coulomb(P)
    assert seen == n * (n + 1) // 2 and torch.equal(acc, full)
    # balanced pair-row blocks through the sharding helpers
    os.environ["GDFT_PACK_ERI"] = "always"
    acc2, acc3 = torch.zeros_like(full), torch.zeros_like(full)
    for r in range(world):
        part = gdist.shard_molecule_tensors({"weights": mol["weights"], "rep_tensor": eri}, r, world, shard_eri="pairs")
        shard = gdist.GridShard(None, r, world, None, part["eri_pair0"])
        acc2 += gdist.local_coulomb(P, part["rep_tensor"], shard)
        os.environ["GDFT_PACK_ERI"] = "never"
        acc3 += gdist.local_coulomb(P, part["rep_tensor"], shard)  # same rows, plain sweep + scatter
        os.environ["GDFT_PACK_ERI"] = "always"
    assert torch.equal(acc2, full)
    assert rel(acc3, full) < 1e-13


def test_autograd_closes_over_the_packed_sweep(cuda_device):
    n = 18
    mol = synthetic_molecule(32, n, seed=5)
    eri = mol["rep_tensor"].to(cuda_device)
    os.environ["GDFT_PACK_ERI"] = "always"
    P = (mol["rdm1"].sum(0).to(cuda_device) + 0.05 * torch.randn(n, n, dtype=F64, device=cuda_device)).requires_grad_(True)
    w = torch.randn(n, n, dtype=F64, device=cuda_device)
    E = (ops.coulomb_j_auto(P, eri) * w).sum() + 0.5 * (P * ops.coulomb_j_auto(P, eri)).sum()
    (g,) = torch.autograd.grad(E, P, create_graph=True)
    Pc = P.detach().cpu().requires_grad_(True)
    Jc = oracle.coulomb_potential(Pc, mol["rep_tensor"])
    Ec = (Jc * w.cpu()).sum() + 0.5 * (Pc * Jc).sum()
    (gc,) = torch.autograd.grad(Ec, Pc, create_graph=True)
    assert rel(g.detach().cpu(), gc.detach()) < 1e-12
    u = torch.randn(n, n, dtype=F64, device=cuda_device)
    (h,) = torch.autograd.grad((g * u).sum(), P)
    (hc,) = torch.autograd.grad((gc * u.cpu()).sum(), Pc)
    assert rel(h.cpu(), hc) < 1e-12
