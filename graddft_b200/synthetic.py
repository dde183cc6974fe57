"""Seeded synthetic inputs of the shapes named in BASELINE.json.

There is no PySCF here, so molecules are replaced by shape-equivalent random tensors: ao with an
exponential envelope spanning ~1e-6..1 (so densities span many decades and the 1e-30 clip branches are
reachable on a masked fraction of rows), idempotent-like density matrices D_s = C_s C_s^T, an ERI tensor
with the 8-fold permutational symmetry, positive quadrature weights.  Seeds 1984 / 1993 are the ones the
reference's own tests use (tests/unit/test_loss.py:36).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

F64 = torch.float64


def synthetic_molecule(N: int, n: int, *, n_omega: int = 0, seed: int = 1984, device="cpu", with_eri: bool = True,
                       with_grad2: bool = True, symmetric_rdm1: bool = True, mask_frac: float = 1e-3,
                       row_chunk: int = 1 << 18) -> Dict[str, torch.Tensor]:
    """Dict of tensors keyed by the reference's Molecule field names (grad_dft/molecule.py:76-102);
    `grad_n_ao2` stands for grad_n_ao[2], `weights`/`coords` for the Grid fields."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)

    def randn(*shape):
        return torch.randn(*shape, generator=g, dtype=F64, device=dev)

    def rand(*shape):
        return torch.rand(*shape, generator=g, dtype=F64, device=dev)

    alpha = 0.2 + 1.8 * rand(n)
    u = 6.0 * rand(N)
    mask = (rand(N) >= mask_frac).to(F64)
    mol: Dict[str, torch.Tensor] = {}
    ao = torch.empty((N, n), dtype=F64, device=dev)
    gao = torch.empty((N, n, 3), dtype=F64, device=dev)
    g2 = torch.empty((N, n, 3), dtype=F64, device=dev) if with_grad2 else None
    chi = torch.empty((N, n_omega, 2, n), dtype=F64, device=dev) if n_omega else None
    for r0 in range(0, N, row_chunk):
        r1 = min(N, r0 + row_chunk)
        env = torch.exp(-u[r0:r1, None] * alpha[None, :]) * mask[r0:r1, None]
        ao[r0:r1] = randn(r1 - r0, n) * env
        gao[r0:r1] = randn(r1 - r0, n, 3) * env[:, :, None]
        if with_grad2:
            g2[r0:r1] = randn(r1 - r0, n, 3) * env[:, :, None]
        if n_omega:
            chi[r0:r1] = randn(r1 - r0, n_omega, 2, n) * env[:, None, None, :]
    mol["ao"], mol["grad_ao"] = ao, gao
    if with_grad2:
        mol["grad_n_ao2"] = g2
    if n_omega:
        mol["chi"] = chi
        mol["omegas"] = torch.tensor([0.0, 0.4, 0.8, 1.2][:n_omega], dtype=F64, device=dev)
    mol["weights"] = rand(N) * (4.0 * math.pi * 36.0 / N)
    mol["coords"] = randn(N, 3)

    nocc = max(1, math.ceil(n / 6))
    Cs, occs = [], []
    for s in range(2):
        q, _ = torch.linalg.qr(randn(n, n))
        Cs.append(q)
        occ = torch.zeros(n, dtype=F64, device=dev)
        occ[:nocc] = 1.0
        occs.append(occ)
    mo_coeff, mo_occ = torch.stack(Cs), torch.stack(occs)
    rdm1 = torch.einsum("sij,sj,skj->sik", mo_coeff, mo_occ, mo_coeff)
    if not symmetric_rdm1:
        rdm1 = rdm1 + 0.05 * randn(2, n, n) / math.sqrt(n)
    mol["rdm1"], mol["mo_coeff"], mol["mo_occ"] = rdm1, mo_coeff, mo_occ
    mol["mo_energy"] = torch.sort(randn(2, n), dim=1).values
    h = randn(n, n)
    mol["h1e"] = 0.5 * (h + h.T)
    sm = randn(n, n) / math.sqrt(n)
    mol["s1e"] = torch.eye(n, dtype=F64, device=dev) + 0.05 * (sm + sm.T)
    mol["nuclear_repulsion"] = torch.tensor(1.2345678901234567, dtype=F64, device=dev)
    if with_eri:
        Q = 2 * n
        B = randn(Q, n, n)
        B = 0.5 * (B + B.transpose(1, 2))
        B2 = B.reshape(Q, n * n)
        mol["rep_tensor"] = ((B2.T @ B2) / Q).reshape(n, n, n, n)
    return mol


def to_device(mol: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    return {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in mol.items()}
