"""The kernels at BASELINE.json's FULL sizes (C4: 2,000,000 grid points x 400 AOs; C3: n = 264 rep_tensor, 38.9 GB),
where the CPU oracle cannot follow, checked through size-independent properties of the path:

  * adjointness   <L D, c> = <D, L^T c>  (the forward and transpose kernels are each other's VJP),
  * linearity     L(a D1 + b D2) = a L D1 + b L D2,
  * restriction   the rows of the full-size result on a sampled block of grid rows equal the result of the same
                  kernel on that block alone, which IS small enough for the oracle -- tying the full-size launch to the
                  reference formulas,
  * additivity    E_xc and V_xc of the full grid = the sum over two halves of the grid (what grid sharding relies on),
  * symmetry      J = J^T for a rep_tensor with the (pq)<->(rt) symmetry, and J rows against the oracle on a row sample.

Tolerances are relative to the largest entry; 1e-12 leaves room for the different summation orders only.
"""
import pytest
import torch

import oracle
import graddft_b200 as gd
from graddft_b200 import ops
from graddft_b200._lib import GDFT_GRAD, GDFT_RHO
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu
F64 = torch.float64
N4, n4 = 2_000_000, 400


def rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))


@pytest.fixture(scope="module")
def c4(cuda_device):
    if torch.cuda.get_device_properties(cuda_device).total_memory < 120e9:
        pytest.skip("needs a 180 GB device")
    mol = synthetic_molecule(N4, n4, seed=1984, device=cuda_device, with_eri=False, with_grad2=False, symmetric_rdm1=False)
    basis = ops.PackedBasis(mol["ao"], mol["grad_ao"])
    yield mol, basis
    del mol, basis
    torch.cuda.empty_cache()


def test_c4_density_adjoint_linear_restriction(cuda_device, c4):
    mol, basis = c4
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(7)
    D1 = mol["rdm1"]
    D2 = torch.randn(2, n4, n4, generator=g, dtype=F64, device=dev) / n4
    flags = GDFT_RHO | GDFT_GRAD
    with torch.no_grad():
        rho1, grho1 = ops.density_forward(basis, D1, flags)[:2]
        rho2, grho2 = ops.density_forward(basis, D2, flags)[:2]
        rho3, grho3 = ops.density_forward(basis, 0.7 * D1 - 1.3 * D2, flags)[:2]
        assert rel(rho3, 0.7 * rho1 - 1.3 * rho2) < 1e-12
        assert rel(grho3, 0.7 * grho1 - 1.3 * grho2) < 1e-12
        # adjointness against the split-K transpose kernel
        cr = torch.randn(N4, 2, generator=g, dtype=F64, device=dev)
        cg = torch.randn(N4, 2, 3, generator=g, dtype=F64, device=dev)
        Dbar = ops.density_transpose(basis, cr, cg)
        lhs = (rho2 * cr).sum() + (grho2 * cg).sum()
        rhs = (D2 * Dbar).sum()
        assert abs(float(lhs - rhs)) < 1e-11 * max(abs(float(lhs)), float((rho2.abs() * cr.abs()).sum()))
        # run-to-run reproducibility of the split-K reduction
        assert torch.equal(Dbar, ops.density_transpose(basis, cr, cg))
        # restriction: a block of rows that starts inside a CTA tile and is not a multiple of it
        lo, hi = 1_234_567, 1_234_567 + 3001
        sub = ops.PackedBasis(mol["ao"][lo:hi].contiguous(), mol["grad_ao"][lo:hi].contiguous())
        rs, gs = ops.density_forward(sub, D1, flags)[:2]
        assert rel(rs, rho1[lo:hi]) < 1e-13 and rel(gs, grho1[lo:hi]) < 1e-13
    ao_c, gao_c, D_c = mol["ao"][lo:hi].cpu(), mol["grad_ao"][lo:hi].cpu(), D1.cpu()
    assert rel(rho1[lo:hi].cpu(), oracle.density(D_c, ao_c)) < 1e-12
    assert rel(grho1[lo:hi].cpu(), oracle.grad_density(D_c, ao_c, gao_c)) < 1e-12
    # the transpose restricted to the same block against the oracle's closed formula (cotangents zero elsewhere)
    with torch.no_grad():
        Dsub = ops.density_transpose(sub, cr[lo:hi].contiguous(), cg[lo:hi].contiguous())
    ref = oracle.density_vjp_formula(ao_c, gao_c, torch.zeros_like(ao_c), cr[lo:hi].cpu(), cg[lo:hi].cpu())
    assert rel(Dsub.cpu(), ref) < 1e-12


def test_c4_xc_build_is_additive_over_the_grid(cuda_device, c4):
    """What grid sharding relies on (one all-reduce of [E_xc | V_xc]): build(full grid) = build(rows < h) + build(rows >= h)."""
    mol, basis = c4
    h = 1_000_064  # a multiple of the 128-row tile, as distributed.shard_bounds cuts
    keys = ("rdm1", "mo_coeff", "mo_occ", "mo_energy", "h1e", "s1e", "nuclear_repulsion")

    def build(lo, hi):
        part = {k: mol[k] for k in keys}
        part.update(ao=mol["ao"][lo:hi], grad_ao=mol["grad_ao"][lo:hi], weights=mol["weights"][lo:hi], coords=mol["coords"][lo:hi])
        m = gd.molecule_from_tensors(part, cuda_device)
        e, v, _ = gd.xc_energy_and_grads(gd.B88, None, m.rdm1, m)
        return e.detach(), v.detach()

    e_full, v_full = build(0, N4)
    e_a, v_a = build(0, h)
    e_b, v_b = build(h, N4)
    assert abs(float(e_full - (e_a + e_b))) < 1e-12 * abs(float(e_full))
    assert rel(v_a + v_b, v_full) < 1e-12
    assert bool(torch.isfinite(v_full).all())


def test_c3_eri_sweep_full_size(cuda_device):
    """n = 264: the 38.9 GB rep_tensor of the benzene shape.  Symmetry, linearity, E_J = <P, J>/2, and J on a sample of
    (p,q) rows against the oracle's einsum."""
    if torch.cuda.get_device_properties(cuda_device).total_memory < 120e9:
        pytest.skip("needs a 180 GB device")
    n = 264
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(11)
    Q = 2 * n
    B = torch.randn(Q, n, n, generator=g, dtype=F64, device=dev)
    B = (0.5 * (B + B.transpose(1, 2))).reshape(Q, n * n)
    eri = ((B.T @ B) / Q).reshape(n, n, n, n)
    del B
    P1 = torch.randn(n, n, generator=g, dtype=F64, device=dev)
    P1 = P1 + P1.T
    P2 = torch.randn(n, n, generator=g, dtype=F64, device=dev)
    with torch.no_grad():
        J1, EJ = ops.coulomb_j_and_energy(P1, eri)
        J2 = ops.coulomb_j(P2, eri)
        J3 = ops.coulomb_j(0.3 * P1 + 2.0 * P2, eri)
    assert rel(J1, J1.T) < 1e-12                       # (pq|rt) = (qp|rt)
    assert rel(J3, 0.3 * J1 + 2.0 * J2) < 1e-12
    assert abs(float(EJ) - 0.5 * float((P1 * J1).sum())) < 1e-11 * abs(float(EJ))
    rows = torch.tensor([0, 1, 263, 264, 12345, 34847, 69695], device=dev)
    # grad_dft/molecule.py:811 on the sampled rows: J[p,q] = sum_rt (pq|rt) P[r,t]
    ref = oracle.coulomb_potential(P2.cpu(), eri.reshape(n * n, n, n)[rows].cpu().reshape(len(rows), 1, n, n)).reshape(-1)
    assert rel(J2.reshape(-1)[rows].cpu(), ref) < 1e-12
    # the packed sweep at full size (9.79 GB of pair rows for the 38.9 GB tensor): symmetry check on the device, the same J from
    # a quarter of the bytes, both triangles from one packed entry, and the policy (packed from the second use on)
    asym, big = ops.eri_symmetry_defect(eri, n)
    assert asym <= 1e-13 * big
    pe = ops.PackedERI.from_rows(eri, n)
    assert pe.complete and pe.packed.numel() * 8 == 9788803200
    with torch.no_grad():
        Jp, EJp = pe.coulomb(P1, want_energy=True)
        Jp2 = pe.coulomb(P2)
    assert rel(Jp, J1) < 1e-13 and rel(Jp2, J2) < 1e-13 and torch.equal(Jp2, Jp2.T)
    assert abs(float(EJp) - float(EJ)) < 1e-12 * abs(float(EJ))
    del pe
    ops.release_packed_eri()
    with torch.no_grad():
        ops.coulomb_j_and_energy(P1, eri)                   # first use of this tensor object: the plain sweep
        Jq, _ = ops.coulomb_j_and_energy(P1, eri)           # second use: checked, packed, swept packed
    assert ops.packed_eri_for(eri, count_use=False) is not None and rel(Jq, J1) < 1e-13
    ops.release_packed_eri()
    del eri
    torch.cuda.empty_cache()


def test_c3_chi_tail_full_chunk(cuda_device):
    """Row f4 at the benzene shape with a 9472-point chunk (4.6 GB of nu, the bench shape): sampled points against the
    oracle's einsum, linearity in rdm1, and both kernels (TMA-fed / register-staged) bit-for-tolerance equal."""
    if torch.cuda.get_device_properties(cuda_device).total_memory < 60e9:
        pytest.skip("needs > 60 GB")
    import os
    from graddft_b200 import interface
    n, rows = 264, 9472
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(5)
    ao = torch.randn(rows, n, generator=g, dtype=F64, device=dev)
    D1 = torch.randn(2, n, n, generator=g, dtype=F64, device=dev)
    D2 = torch.randn(2, n, n, generator=g, dtype=F64, device=dev)
    nu = torch.randn(rows, n, n, generator=g, dtype=F64, device=dev)
    coords = torch.zeros(rows, 3, dtype=F64, device=dev)
    gen = lambda D: interface.generate_chi_tensor(D, ao, coords, lambda c, o: nu, [0.0], chunk_size=None)  # noqa: E731
    old = os.environ.get("GDFT_CHI_TMA")
    try:
        os.environ["GDFT_CHI_TMA"] = "1"
        chi1, chi2, chi3 = gen(D1), gen(D2), gen(0.5 * D1 - 2.0 * D2)
        os.environ["GDFT_CHI_TMA"] = "0"
        chi1_reg = gen(D1)
    finally:
        if old is None:
            os.environ.pop("GDFT_CHI_TMA", None)
        else:
            os.environ["GDFT_CHI_TMA"] = old
    assert chi1.shape == (rows, 1, 2, n)
    assert rel(chi3, 0.5 * chi1 - 2.0 * chi2) < 1e-12
    assert rel(chi1_reg, chi1) < 1e-13
    idx = torch.tensor([0, 7, 8, 4095, 4736, 9463, 9471], device=dev)
    ref = oracle.generate_chi_tensor(D1.cpu(), ao[idx].cpu(), coords[idx].cpu(), lambda c, o: nu[idx].cpu(), [0.0], None)
    assert rel(chi1[idx].cpu(), ref) < 1e-12
