"""From-scratch H2 / STO-3G molecule (TEST INFRASTRUCTURE): the one physically meaningful input of
the test-suite that does not need PySCF.

What `grad_dft/interface/pyscf.py:105-201` (`molecule_from_pyscf`) obtains from PySCF/libcint -- ao, grad_ao, the second
derivatives, h1e, s1e, rep_tensor, the grid, the exact-exchange nu integrals -- is available in closed form for
contracted s-type Gaussians (Boys F0 = erf), so this module builds a real `Molecule`-shaped dict with n = 2 whose
quantities are independently checkable:

* one- and two-electron integrals against Szabo & Ostlund, "Modern Quantum Chemistry", section 3.5.2 (R = 1.4 bohr,
  zeta = 1.24): S12 = 0.6593, T11 = 0.7600, T12 = 0.2365, V11(one centre) = -1.2266, (11|11) = 0.7746,
  (11|22) = 0.5697, (21|11) = 0.4441, (21|21) = 0.2970, E_RHF = -1.1167 Ha;
* grid identities: sum_r w rho = 2, sum_r w tau = <T>, sum_r w lapl(rho) = 0, sum_r w e_HF = -1/2 Tr(D K[D]),
  Tr(D V_x) = 4/3 E_x for Dirac exchange.

The grid is Becke's scheme (J. Chem. Phys. 88, 2547 (1988)): Gauss-Chebyshev (second kind) radial quadrature mapped by
r = R_m (1 + x)/(1 - x), a Gauss-Legendre x uniform-phi product rule on the sphere, fuzzy-cell weights with three
iterations of the smoothing polynomial.  Pure NumPy; field names follow grad_dft/molecule.py:76-102 and
graddft_b200.synthetic.synthetic_molecule.
"""
from __future__ import annotations

import math
from typing import Dict, Sequence

import numpy as np
import torch

# STO-3G hydrogen 1s, zeta = 1.24 (Hehre, Stewart, Pople, J. Chem. Phys. 51, 2657 (1969))
STO3G_H_EXP = np.array([3.42525091, 0.62391373, 0.16885540])
STO3G_H_COEF = np.array([0.15432897, 0.53532814, 0.44463454])

SZABO_OSTLUND = {"S12": 0.6593, "T11": 0.7600, "T12": 0.2365, "V11_one_centre": -1.2266, "V12_one_centre": -0.5974, "V22_at_centre1": -0.6538,
                 "1111": 0.7746, "1122": 0.5697, "2111": 0.4441, "2121": 0.2970, "E_RHF": -1.1167}


def boys_f0(t: np.ndarray) -> np.ndarray:
    t = np.asarray(t, dtype=np.float64)
    small = t < 1e-6
    ts = np.where(small, 1.0, t)
    big = 0.5 * np.sqrt(np.pi / ts) * np.vectorize(math.erf)(np.sqrt(ts))
    return np.where(small, 1.0 - t / 3.0 + t * t / 10.0, big)


def _prims(centers: np.ndarray):
    """Flat list of normalised primitives: (basis index, centre, exponent, coefficient * norm)."""
    out = []
    for a, A in enumerate(centers):
        for al, d in zip(STO3G_H_EXP, STO3G_H_COEF):
            out.append((a, A, al, d * (2.0 * al / math.pi) ** 0.75))
    return out


def one_electron(centers: np.ndarray, charges: Sequence[float]):
    """Overlap, kinetic and nuclear-attraction matrices (the latter per nucleus: V[c, a, b])."""
    n = len(centers)
    S, T, V = np.zeros((n, n)), np.zeros((n, n)), np.zeros((len(charges), n, n))
    P = _prims(centers)
    for (a, A, al, ca) in P:
        for (b, B, be, cb) in P:
            p = al + be
            ab2 = float(((A - B) ** 2).sum())
            K = math.exp(-al * be / p * ab2)
            s = (math.pi / p) ** 1.5 * K
            S[a, b] += ca * cb * s
            T[a, b] += ca * cb * (al * be / p) * (3.0 - 2.0 * al * be / p * ab2) * s
            Pc = (al * A + be * B) / p
            for c, (C, Z) in enumerate(zip(centers, charges)):
                V[c, a, b] += ca * cb * (-Z) * (2.0 * math.pi / p) * K * float(boys_f0(p * ((Pc - C) ** 2).sum()))
    return S, T, V


def two_electron(centers: np.ndarray) -> np.ndarray:
    """(ab|cd) in chemists' notation, the layout of Molecule.rep_tensor (grad_dft/molecule.py:811 contracts the last two)."""
    n = len(centers)
    P = _prims(centers)
    pairs = []
    for (a, A, al, ca) in P:
        for (b, B, be, cb) in P:
            p = al + be
            pairs.append((a, b, p, (al * A + be * B) / p, ca * cb * math.exp(-al * be / p * float(((A - B) ** 2).sum()))))
    eri = np.zeros((n, n, n, n))
    for (a, b, p, Pc, kab) in pairs:
        for (c, d, q, Qc, kcd) in pairs:
            t = p * q / (p + q) * float(((Pc - Qc) ** 2).sum())
            eri[a, b, c, d] += kab * kcd * 2.0 * math.pi ** 2.5 / (p * q * math.sqrt(p + q)) * float(boys_f0(t))
    return eri


def becke_grid(centers: np.ndarray, n_rad: int = 64, n_theta: int = 32, n_phi: int = 8, r_m: float = 1.0):
    """coords[N,3], weights[N] of Becke's multi-centre quadrature (no atomic-size adjustment: homonuclear use)."""
    i = np.arange(1, n_rad + 1)
    x = np.cos(i * math.pi / (n_rad + 1))
    wx = math.pi / (n_rad + 1) * np.sin(i * math.pi / (n_rad + 1))  # Chebyshev-2 weights with the 1/sqrt(1-x^2) folded in
    r = r_m * (1.0 + x) / (1.0 - x)
    wr = wx * 2.0 * r_m / (1.0 - x) ** 2 * r * r
    ct, wt = np.polynomial.legendre.leggauss(n_theta)
    st = np.sqrt(1.0 - ct * ct)
    phi = (np.arange(n_phi) + 0.5) * 2.0 * math.pi / n_phi
    ux = (st[:, None] * np.cos(phi)[None, :]).ravel()
    uy = (st[:, None] * np.sin(phi)[None, :]).ravel()
    uz = np.repeat(ct, n_phi)
    wang = np.repeat(wt, n_phi) * 2.0 * math.pi / n_phi
    unit = np.stack([ux, uy, uz], axis=1)
    coords, weights = [], []
    for A in range(len(centers)):
        pts = centers[A][None, None, :] + r[:, None, None] * unit[None, :, :]
        w = (wr[:, None] * wang[None, :]).ravel()
        pts = pts.reshape(-1, 3)
        dist = np.linalg.norm(pts[:, None, :] - centers[None, :, :], axis=2)  # [pts, atoms]
        cell = np.ones((pts.shape[0], len(centers)))
        for a in range(len(centers)):
            for b in range(len(centers)):
                if a == b:
                    continue
                mu = (dist[:, a] - dist[:, b]) / np.linalg.norm(centers[a] - centers[b])
                for _ in range(3):
                    mu = 1.5 * mu - 0.5 * mu ** 3
                cell[:, a] *= 0.5 * (1.0 - mu)
        coords.append(pts)
        weights.append(w * cell[:, A] / cell.sum(axis=1))
    return np.concatenate(coords), np.concatenate(weights)


def eval_ao(centers: np.ndarray, coords: np.ndarray):
    """ao[N,n], grad_ao[N,n,3], second derivatives d2_i ao [N,n,3] (the diagonal ones: what lapl_density sums,
    grad_dft/molecule.py:474)."""
    N, n = coords.shape[0], len(centers)
    ao, gao, g2 = np.zeros((N, n)), np.zeros((N, n, 3)), np.zeros((N, n, 3))
    for (a, A, al, c) in _prims(centers):
        d = coords - A[None, :]
        e = c * np.exp(-al * (d * d).sum(axis=1))
        ao[:, a] += e
        gao[:, a, :] += (-2.0 * al) * d * e[:, None]
        g2[:, a, :] += (4.0 * al * al * d * d - 2.0 * al) * e[:, None]
    return ao, gao, g2


def nu_integrals(centers: np.ndarray, coords: np.ndarray, omega: float) -> np.ndarray:
    """nu[r, d, a] = int phi_d(r') phi_a(r') f(|r - r'|) dr' with f = 1/r (omega = 0) or erf(omega r)/r: what libcint's
    int1e_grids yields in `_nu_chunk` (grad_dft/external/_hf_density.py:69-103)."""
    N, n = coords.shape[0], len(centers)
    nu = np.zeros((N, n, n))
    P = _prims(centers)
    for (a, A, al, ca) in P:
        for (b, B, be, cb) in P:
            p = al + be
            K = ca * cb * math.exp(-al * be / p * float(((A - B) ** 2).sum()))
            Pc = (al * A + be * B) / p
            r2 = ((coords - Pc[None, :]) ** 2).sum(axis=1)
            if omega == 0:
                nu[:, a, b] += K * (2.0 * math.pi / p) * boys_f0(p * r2)
            else:
                th = omega * omega / (p + omega * omega)
                nu[:, a, b] += K * (2.0 * math.pi / p) * math.sqrt(th) * boys_f0(p * th * r2)
    return nu


def build_h2(R: float = 1.4, n_rad: int = 64, n_theta: int = 32, n_phi: int = 8, omegas: Sequence[float] = (0.0, 0.4)) -> Dict[str, object]:
    """Dict of float64 torch-CPU tensors keyed like `synthetic_molecule`, plus `expected`: closed-form reference values."""
    centers = np.array([[0.0, 0.0, -0.5 * R], [0.0, 0.0, 0.5 * R]])
    charges = [1.0, 1.0]
    S, T, V = one_electron(centers, charges)
    eri = two_electron(centers)
    h = T + V.sum(axis=0)
    c = np.array([1.0, 1.0]) / math.sqrt(S[0, 0] + S[1, 1] + 2.0 * S[0, 1])  # sigma_g, fixed by symmetry
    D = np.outer(c, c)
    rdm1 = np.stack([D, D])
    coords, weights = becke_grid(centers, n_rad, n_theta, n_phi)
    ao, gao, g2 = eval_ao(centers, coords)
    chi = np.stack([np.einsum("sbd,rb,rda->rsa", rdm1, ao, nu_integrals(centers, coords, w)) for w in omegas], axis=1)

    Ptot = 2.0 * D
    J = np.einsum("pqrt,rt->pq", eri, Ptot)
    Kx = np.einsum("prqt,rt->pq", eri, D)  # per spin
    e_nuc = 1.0 / R
    e1 = float((Ptot * h).sum())
    ej = 0.5 * float((Ptot * J).sum())
    ex = -0.5 * 2.0 * float((D * Kx).sum())
    fock = h + J - Kx
    eps = float(c @ fock @ c)
    mo_b = np.array([1.0, -1.0]) / math.sqrt(S[0, 0] + S[1, 1] - 2.0 * S[0, 1])
    t = torch.from_numpy
    mol: Dict[str, object] = {
        "ao": t(ao), "grad_ao": t(gao), "grad_n_ao2": t(g2), "chi": t(np.ascontiguousarray(chi)), "omegas": t(np.asarray(omegas, dtype=np.float64)),
        "weights": t(weights), "coords": t(coords), "rdm1": t(rdm1), "h1e": t(h), "s1e": t(S), "rep_tensor": t(eri),
        "nuclear_repulsion": torch.tensor(e_nuc, dtype=torch.float64),
        "mo_coeff": t(np.stack([np.stack([c, mo_b], axis=1)] * 2)), "mo_occ": t(np.array([[1.0, 0.0], [1.0, 0.0]])),
        "mo_energy": t(np.array([[eps, float(mo_b @ fock @ mo_b)]] * 2)),
    }
    mol["expected"] = {"S": S, "T": T, "V_per_nucleus": V, "eri": eri, "E_nuc": e_nuc, "E_1": e1, "E_J": ej, "E_x_HF": ex,
                       "E_RHF": e_nuc + e1 + ej + ex, "kinetic": float((Ptot * T).sum()), "nonXC": e_nuc + e1 + ej, "electrons": 2.0}
    return mol
