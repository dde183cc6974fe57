"""Row f4 on the B200: gdft_chi_contract / generate_chi_tensor against the golden chi of the reference's own
generate_chi_tensor (tests/golden/io_chi.npz) and against the oracle at the shapes the kernel specialises on."""
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
import graddft_b200 as gd
from graddft_b200 import interface, ops
from test_interface_cpu import seeded_nu

pytestmark = pytest.mark.gpu
G = Path(__file__).resolve().parent / "golden"
F64 = torch.float64


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("tag", ["a", "b"])  # a: n = 10 (128-bit path), b: n = 7 (odd n: scalar path)
@pytest.mark.parametrize("where", ["device", "host", "pinned", "pinned_reused"])
def test_generate_chi_tensor_golden(cuda_device, tag, where):
    z = np.load(G / "io_chi.npz")
    t = lambda k: torch.from_numpy(z[f"{tag}_{k}"])  # noqa: E731
    n = t("ao").shape[1]
    provider = seeded_nu(n, int(z[f"{tag}_nu_seed"]))
    calls = []
    reused = torch.empty((int(z[f"{tag}_chunk"]), n, n), dtype=F64).pin_memory()

    def nu_fn(coords, omega):
        calls.append(len(coords))
        nu = provider(coords, omega)
        if where == "device":
            return nu.to(cuda_device)
        if where == "pinned_reused":  # a provider that writes every chunk into the SAME page-locked buffer
            reused[:len(coords)].copy_(nu)
            return reused[:len(coords)]
        return nu.contiguous().pin_memory() if where == "pinned" else nu.numpy()

    chunk = int(z[f"{tag}_chunk"])
    chi = interface.generate_chi_tensor(t("rdm1").to(cuda_device), t("ao").to(cuda_device), t("coords").to(cuda_device), nu_fn,
                                        [float(o) for o in z[f"{tag}_omegas"]], chunk_size=chunk)
    ref = t("out_chi")
    assert chi.shape == ref.shape and chi.is_cuda
    assert rel(chi.cpu(), ref) < 1e-13
    N = ref.shape[0]
    assert calls == [min(chunk, N - s) for _ in range(2) for s in range(0, N, chunk)]


def test_generate_chi_tensor_edge_cases(cuda_device):
    ao = torch.randn(9, 4, dtype=F64, device=cuda_device)
    D = torch.randn(2, 4, 4, dtype=F64, device=cuda_device)
    nu = lambda c, o: torch.zeros(len(c), 4, 4, dtype=F64, device=cuda_device)  # noqa: E731
    assert interface.generate_chi_tensor(D, ao, ao[:, :3], nu, []).numel() == 0
    with pytest.raises(ValueError):
        interface.generate_chi_tensor(D, ao, ao[:, :3], nu, [0.0, -0.4])
    with pytest.raises(TypeError):
        interface.generate_chi_tensor(D, ao, ao[:, :3], lambda c, o: torch.zeros(len(c), 3, 4, dtype=F64, device=cuda_device), [0.0])
    chi = interface.generate_chi_tensor(D, ao, ao[:, :3], nu, [0.0], chunk_size=None)
    assert chi.shape == (9, 1, 2, 4) and float(chi.abs().max()) == 0.0


# one pass narrow / two rows in flight (NJ <= 5) / one row in flight (NJ 6..8) / two passes (n > 512) / odd n; ragged chunks
@pytest.mark.parametrize("N,n,chunk", [(37, 2, 8), (300, 43, 128), (257, 264, 100), (130, 400, 64), (70, 520, 33), (41, 129, 16), (9, 1, 4)])
@pytest.mark.parametrize("kernel", ["auto", "0", "1"])  # GDFT_CHI_TMA: heuristic / register-staged / TMA-fed
def test_chi_contract_vs_oracle(cuda_device, N, n, chunk, kernel, monkeypatch):
    if kernel == "auto":
        monkeypatch.delenv("GDFT_CHI_TMA", raising=False)
    else:
        monkeypatch.setenv("GDFT_CHI_TMA", kernel)
    g = torch.Generator().manual_seed(1984 + n)
    ao = torch.randn(N, n, generator=g, dtype=F64)
    D = torch.randn(2, n, n, generator=g, dtype=F64)  # non-symmetric: pins the index placement of "...bd,b,da->...a"
    nus = torch.randn(2, N, n, n, generator=g, dtype=F64)
    omegas = [0.0, 0.4]
    coords = torch.arange(N, dtype=F64)[:, None].expand(N, 3)

    def nu_cpu(c, omega):
        i = int(c[0, 0])
        return nus[omegas.index(omega), i:i + len(c)]

    ref = oracle.generate_chi_tensor(D, ao, coords, nu_cpu, omegas, chunk)
    nus_d = nus.to(cuda_device)
    chi = interface.generate_chi_tensor(D.to(cuda_device), ao.to(cuda_device), coords.to(cuda_device),
                                        lambda c, omega: nus_d[omegas.index(omega), int(c[0, 0]):int(c[0, 0]) + len(c)], omegas, chunk)
    assert rel(chi.cpu(), ref) < 1e-13
    # chi generated this way feeds the hot path: HF energy density through the kernels == through the oracle
    if n >= 2:
        m = gd.Molecule(grid=gd.Grid(coords.to(cuda_device), torch.ones(N, dtype=F64, device=cuda_device)), atom_index=None, nuclear_pos=None,
                        ao=ao.to(cuda_device), grad_ao=None, grad_n_ao=None, rdm1=D.to(cuda_device), nuclear_repulsion=None, h1e=None, vj=None,
                        mo_coeff=None, mo_occ=None, mo_energy=None, omegas=torch.tensor(omegas, dtype=F64), chi=chi)
        ehf = m.HF_energy_density(omegas)
        assert rel(ehf.cpu(), oracle.HF_energy_density(D, ao, ref)) < 1e-12


def test_chi_contract_status_codes(cuda_device):
    L = ops.lib()
    z = torch.zeros(8, dtype=F64, device=cuda_device)
    s = ops.stream_ptr()
    assert L.gdft_chi_contract(s, 0, 2, ops.ptr(z), 2, ops.ptr(z), ops.ptr(z), ops.ptr(z), 4) == 1
    assert L.gdft_chi_contract(s, 1, 2, ops.ptr(z), 1, ops.ptr(z), ops.ptr(z), ops.ptr(z), 4) == 1   # ao_ld < n
    assert L.gdft_chi_contract(s, 1, 2, ops.ptr(z), 2, ops.ptr(z), ops.ptr(z), ops.ptr(z), 3) == 1   # chi_ld < 2n
    assert L.gdft_chi_contract(s, 1, 2, None, 2, ops.ptr(z), ops.ptr(z), ops.ptr(z), 4) == 5
    assert L.gdft_chi_contract(s, 1, int(L.gdft_chi_contract_max_n()) + 1, ops.ptr(z), 4096, ops.ptr(z), ops.ptr(z), ops.ptr(z), 8192) == 1
