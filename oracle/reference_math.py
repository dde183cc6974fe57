"""float64 torch-CPU restatement of the Grad DFT hot-path arithmetic (TEST INFRASTRUCTURE).

Every function names the reference lines it restates (paths relative to /root/reference).
Tensors are plain ``torch.Tensor`` (float64, CPU); autograd through these expressions provides the
reference VJPs.  jnp -> torch conventions used throughout:
  jnp.clip(x, a_min=c)   -> torch.clamp(x, min=c)     (zero gradient where clipped)
  jnp.where(c, a, b)     -> torch.where(c, a, b)      (gradient only through the selected branch)
  2 ** x                 -> torch.exp2(x)
"""
from __future__ import annotations

import math
from typing import Callable, Optional, Sequence

import torch

CLIP = 1e-30
F64 = torch.float64

__all__ = [
    "CLIP", "abs_clip", "density", "grad_density", "lapl_density", "kinetic_density",
    "HF_energy_density", "HF_fock", "FactorizedERI", "coulomb_potential", "coulomb_energy", "one_body_energy",
    "nonXC", "make_rdm1", "orbital_grad", "get_occ", "integrate", "xc_energy",
    "exchange_polarization_correction", "correlation_polarization_correction",
    "lsda_x_e", "b88_x_e", "pw92_c_e", "vwn_c_e", "lyp_c_e",
    "b3lyp_exhf_densities", "b3lyp_combine", "B3LYP_COEFFS",
    "dm21_coefficient_inputs", "dm21_densities", "dm21_combine_cinputs", "dm21_combine_densities",
    "dm21_mlp", "dm21_mlp_init", "mgga_feature_densities",
    "xc_energy_of_rdm1", "predict_b3lyp", "predict_semilocal", "predict_dm21",
    "density_vjp_formula", "safe_fock_solver", "jittable_diis_run", "diff_scf_loop_energy",
    "predict_dm21_traced", "diff_simple_scf_loop_energy", "generate_chi_tensor",
]


# --------------------------------------------------------------------------------------------
# grid / AO tensor algebra  (grad_dft/molecule.py)
# --------------------------------------------------------------------------------------------
def abs_clip(x: torch.Tensor, threshold: float = CLIP) -> torch.Tensor:
    """molecule.py:687-689 -- zero everything with |x| <= threshold (value and gradient)."""
    return torch.where(x.abs() > threshold, x, torch.zeros_like(x))


def density(rdm1, ao):
    """molecule.py:409 -- rho[r,s] = sum_ab D[s,a,b] ao[r,a] ao[r,b]."""
    return torch.einsum("sab,ra,rb->rs", rdm1, ao, ao)


def grad_density(rdm1, ao, grad_ao):
    """molecule.py:440 -- 2 * sum_ab D[s,a,b] ao[r,a] grad_ao[r,b,j] (a on ao, b on grad_ao)."""
    return 2.0 * torch.einsum("sab,ra,rbj->rsj", rdm1, ao, grad_ao)


def lapl_density(rdm1, ao, grad_ao, grad_2_ao):
    """molecule.py:472-474 -- 2 sum D dao.dao + 2 sum_ab D ao[r,a] (sum_i d2ao[r,b,i])."""
    return 2.0 * torch.einsum("sab,raj,rbj->rs", rdm1, grad_ao, grad_ao) + 2.0 * torch.einsum(
        "sab,ra,rbi->rs", rdm1, ao, grad_2_ao
    )


def kinetic_density(rdm1, grad_ao):
    """molecule.py:502 -- tau[r,s] = 1/2 sum_ab D[s,a,b] dao[r,a,j] dao[r,b,j]."""
    return 0.5 * torch.einsum("sab,raj,rbj->rs", rdm1, grad_ao, grad_ao)


def HF_energy_density(rdm1, ao, chi):
    """molecule.py:537-541 -- e_HF[w,s,r] = -1/2 sum_{a,c} chi[r,w,s,c] D[s,a,c] ao[r,a]."""
    return -0.5 * torch.einsum("rwsc,sac,ra->wsr", chi, rdm1, ao)


def HF_fock(chi, g, ao):
    """molecule.py:606-613 (and 678-685) -- F[w,s,a,c] = -1/2 sum_r chi[r,w,s,c] g[w,s,r] ao[r,a].

    The reference maps over the orbital index c of chi.transpose(3,0,1,2), producing [c,w,s,a], then
    transposes (1,2,3,0) -> [w,s,a,c]."""
    return -0.5 * torch.einsum("rwsc,wsr,ra->wsac", chi, g, ao)


class FactorizedERI:
    """Test-only stand-in for a rep_tensor given in factorised form, (pq|rt) = sum_Q B[Q,p,q] B[Q,r,t] / scale (how the
    synthetic tensors are built): the contraction of molecule.py:811 without materialising n^4 doubles on the host, so
    that predictor-level parity can be checked at n = 264 (38.9 GB dense).  `dense(device)` is the tensor itself."""

    def __init__(self, B: torch.Tensor, scale: float):
        self.B, self.scale = B, float(scale)

    def contract(self, P: torch.Tensor) -> torch.Tensor:
        return torch.einsum("Qpq,Q->pq", self.B, torch.einsum("Qrt,rt->Q", self.B, P)) / self.scale

    def dense(self, device=None) -> torch.Tensor:
        B = self.B.to(device) if device is not None else self.B
        Q, n = B.shape[0], B.shape[1]
        B2 = B.reshape(Q, n * n)
        return ((B2.T @ B2) / self.scale).reshape(n, n, n, n)


def coulomb_potential(P, rep_tensor):
    """molecule.py:811 -- J[p,q] = sum_rt (pq|rt) P[r,t]."""
    if isinstance(rep_tensor, FactorizedERI):
        return rep_tensor.contract(P)
    return torch.einsum("pqrt,rt->pq", rep_tensor, P)


def coulomb_energy(P, rep_tensor):
    """molecule.py:781-783 -- E_J = 1/2 <P, J>."""
    return 0.5 * torch.einsum("pq,pq->", P, coulomb_potential(P, rep_tensor))


def one_body_energy(P, h1e):
    """molecule.py:756."""
    return torch.einsum("ij,ij->", P, h1e)


def nonXC(P, h1e, rep_tensor, nuclear_repulsion):
    """molecule.py:727-733 -- E_nuc + E_1 + E_J for the spin-summed density matrix P."""
    return nuclear_repulsion + one_body_energy(P, h1e) + coulomb_energy(P, rep_tensor)


def make_rdm1(mo_coeff, mo_occ):
    """molecule.py:846 -- D[s,i,k] = sum_j C[s,i,j] occ[s,j] C[s,k,j]."""
    return torch.einsum("sij,sj,skj->sik", mo_coeff, mo_occ, mo_coeff)


def orbital_grad(mo_coeff, mo_occ, F):
    """molecule.py:378-381 -- C_vir^T F C_occ summed over spin, with zero-masked (not sliced) blocks."""
    occ = (mo_occ > 0).unsqueeze(1)
    vir = (mo_occ == 0).unsqueeze(1)
    C_occ = torch.where(occ, mo_coeff, torch.zeros_like(mo_coeff))
    C_vir = torch.where(vir, mo_coeff, torch.zeros_like(mo_coeff))
    return torch.einsum("sab,sac,scd->bd", C_vir, F, C_occ)


def get_occ(mo_energy, nelecs):
    """molecule.py:851-889 -- aufbau: the nelec[s] lowest orbitals of each spin get occupation 1."""
    occ = torch.zeros_like(mo_energy)
    for s in range(2):
        order = torch.argsort(mo_energy[s], stable=True)
        occ[s, order[: int(nelecs[s])]] = 1.0
    return occ


# --------------------------------------------------------------------------------------------
# XC integral  (grad_dft/functional.py)
# --------------------------------------------------------------------------------------------
def integrate(energy_density, weights, clip: float = CLIP):
    """functional.py:342 -- sum_r aclip(w_r) aclip(e_r)."""
    return torch.einsum("r,r->", abs_clip(weights, clip), abs_clip(energy_density, clip))


def xc_energy(coefficients, densities, weights, clip: float = CLIP):
    """functional.py:250-253 -- e_r = sum_f c[r,f] d[r,f]; aclip; quadrature.  ``coefficients`` may be
    [1,F] (constant functionals, popular_functionals.py:347) and is broadcast like jnp.einsum does."""
    if coefficients.shape[0] == 1 and densities.shape[0] != 1:
        coefficients = coefficients.expand(densities.shape[0], -1)
    e = torch.einsum("rf,rf->r", coefficients, densities)
    return integrate(abs_clip(e, clip), weights, clip)


# --------------------------------------------------------------------------------------------
# spin interpolation  (grad_dft/functional.py:950-1045)
# --------------------------------------------------------------------------------------------
_FZ_DEN = 2.0 * (2.0 ** (1.0 / 3.0) - 1.0)


def exchange_polarization_correction(e_PF, rho):
    """functional.py:973-979 (zeta unguarded; plain powers)."""
    zeta = (rho[:, 0] - rho[:, 1]) / rho.sum(dim=1)
    fz = ((1 - zeta) ** (4.0 / 3.0) + (1 + zeta) ** (4.0 / 3.0) - 2.0) / _FZ_DEN
    return e_PF[:, 0] + (e_PF[:, 1] - e_PF[:, 0]) * fz


def _fzeta_log(z):
    """functional.py:1014-1017 (log2/exp2 domain)."""
    zm = torch.exp2(4.0 * torch.log2(1 - z) / 3.0)
    zp = torch.exp2(4.0 * torch.log2(1 + z) / 3.0)
    return (zm + zp - 2.0) / _FZ_DEN


def _fzeta_pp0() -> float:
    """functional.py:1040 -- grad(grad(fzeta))(0.) evaluated by autodiff on the log-domain form."""
    z = torch.zeros((), dtype=F64, requires_grad=True)
    (g1,) = torch.autograd.grad(_fzeta_log(z), z, create_graph=True)
    (g2,) = torch.autograd.grad(g1, z)
    return float(g2)


FZETA_PP0 = _fzeta_pp0()


def _pw_G(log_rs, A, a1, b1, b2, b3, b4):
    """The PW92 G-function in the reference's exp2/log2 form (functional.py:1026-1032,
    popular_functionals.py:127-135); returns 2A(1+a1 rs) ln(1 + 1/(2A(...))) (positive)."""
    ars = torch.exp2(math.log2(a1) + log_rs)
    brs_1_2 = torch.exp2(math.log2(b1) + log_rs / 2)
    brs = torch.exp2(math.log2(b2) + log_rs)
    brs_3_2 = torch.exp2(math.log2(b3) + 3 * log_rs / 2)
    brs2 = torch.exp2(math.log2(b4) + 2 * log_rs)
    return 2 * A * (1 + ars) * torch.log(1 + (1 / (2 * A)) / (brs_1_2 + brs + brs_3_2 + brs2))


_LOG2_RS0 = math.log2((3.0 / (4.0 * math.pi)) ** (1.0 / 3.0))


def correlation_polarization_correction(e_PF, rho, clip: float = CLIP):
    """functional.py:1008-1045."""
    rho_t = rho.sum(dim=1)
    log_rho = torch.log2(torch.clamp(rho_t, min=clip))
    log_rs = _LOG2_RS0 - log_rho / 3.0
    zeta = torch.where(rho_t > clip, (rho[:, 0] - rho[:, 1]) / rho_t, torch.zeros_like(rho_t))
    alphac = _pw_G(log_rs, 0.016887, 0.11125, 10.357, 3.6231, 0.88026, 0.49671)
    fz = _fzeta_log(zeta)
    z4 = zeta ** 4
    return e_PF[:, 0] + alphac * (fz / FZETA_PP0) * (1 - z4) + (e_PF[:, 1] - e_PF[:, 0]) * fz * z4


# --------------------------------------------------------------------------------------------
# closed-form energy densities  (grad_dft/popular_functionals.py)
# --------------------------------------------------------------------------------------------
def lsda_x_e(rho, clip: float = CLIP):
    """popular_functionals.py:41-50."""
    rho = torch.clamp(rho, min=clip)
    pref = -0.75 * (torch.tensor([[3.0, 6.0]], dtype=rho.dtype) / math.pi) ** (1.0 / 3.0)
    lda_es = pref * rho.sum(dim=1, keepdim=True) ** (4.0 / 3.0)
    return exchange_polarization_correction(lda_es, rho)


def b88_x_e(rho, grad_rho, clip: float = CLIP):
    """popular_functionals.py:70-103."""
    beta = 0.0042
    rho = torch.clamp(rho, min=clip)
    log_rho = torch.log2(torch.clamp(rho, min=clip))
    sigma = (grad_rho ** 2).sum(dim=-1)
    log_g = torch.log2(torch.clamp(sigma, min=clip)) / 2
    log_x = log_g - 4.0 / 3.0 * log_rho
    x = torch.exp2(log_x)
    e = beta * torch.exp2(4 * log_rho / 3 + 2 * log_x - torch.log2(1 + 6 * beta * x * torch.asinh(x)))
    return -e.sum(dim=1)


def pw92_c_e(rho, clip: float = CLIP):
    """popular_functionals.py:120-139 (rho is NOT pre-clipped here)."""
    rho_t = rho.sum(dim=1, keepdim=True)
    log_rho = torch.log2(torch.clamp(rho_t, min=clip))
    log_rs = _LOG2_RS0 - log_rho / 3.0
    eP = -_pw_G(log_rs, 0.031091, 0.21370, 7.5957, 3.5876, 1.6382, 0.49294)
    eF = -_pw_G(log_rs, 0.015545, 0.20548, 14.1189, 6.1977, 3.3662, 0.62517)
    e_PF = torch.cat([eP, eF], dim=1)
    return correlation_polarization_correction(e_PF, rho, clip) * rho.sum(dim=1)


def vwn_c_e(rho, clip: float = CLIP):
    """popular_functionals.py:158-195."""
    A = torch.tensor([[0.0621814, 0.0621814 / 2]], dtype=rho.dtype)
    b = torch.tensor([[3.72744, 7.06042]], dtype=rho.dtype)
    c = torch.tensor([[12.9352, 18.0578]], dtype=rho.dtype)
    x0 = torch.tensor([[-0.10498, -0.325]], dtype=rho.dtype)
    rho = torch.where(rho > clip, rho, torch.zeros_like(rho))
    log_rho = torch.log2(torch.clamp(rho.sum(dim=1, keepdim=True), min=clip))
    log_rs = _LOG2_RS0 - log_rho / 3.0
    log_x = log_rs / 2
    x = torch.exp2(log_x)
    X = torch.exp2(2 * log_x) + torch.exp2(log_x + torch.log2(b)) + c
    X0 = x0 ** 2 + b * x0 + c
    Q = torch.sqrt(4 * c - b ** 2)
    at = torch.atan(Q / (2 * x + b))
    e_PF = A / 2 * (
        2 * torch.log(x) - torch.log(X) + 2 * b / Q * at
        - b * x0 / X0 * (torch.log((x - x0) ** 2 / X) + 2 * (2 * x0 + b) / Q * at)
    )
    return correlation_polarization_correction(e_PF, rho, clip) * rho.sum(dim=1)


def lyp_c_e(rho, grad_rho, grad2rho, clip: float = CLIP):
    """popular_functionals.py:229-269 (original 1988 LYP with the Laplacian)."""
    a, b, c, d = 0.04918, 0.132, 0.2533, 0.349
    CF = 0.3 * (3 * math.pi ** 2) ** (2.0 / 3.0)
    rho = torch.clamp(rho, min=clip)
    sigma = (grad_rho ** 2).sum(dim=-1)
    zero = torch.zeros_like(rho)
    t = (torch.where(rho > clip, sigma / rho, zero) - grad2rho) / 8.0
    rho_t = rho.sum(dim=1)
    frac = torch.where(rho_t > clip, (rho ** 2).sum(dim=1) / rho_t ** 2, torch.ones_like(rho_t))
    gamma = 2 * (1 - frac)
    rhos_ts = rho_t * t.sum(dim=1)
    rho_tw = (rho * t).sum(dim=1)
    rho_lap = (rho * grad2rho).sum(dim=1)
    rhom1_3 = rho_t ** (-1.0 / 3.0)
    rho8_3 = (rho ** (8.0 / 3.0)).sum(dim=1)
    rhom5_3 = rho_t ** (-5.0 / 3.0)
    expf = torch.where(rho_t > 0, torch.exp(-c * rhom1_3), torch.zeros_like(rho_t))
    par = 2 ** (2.0 / 3.0) * CF * rho8_3 - rhos_ts + rho_tw / 9 + rho_lap / 18
    brk = torch.where(rho_t > clip, 2 * b * rhom5_3 * par * expf, torch.zeros_like(rho_t))
    return -a * torch.where(rho_t > clip, gamma / (1 + d * rhom1_3) * (rho_t + brk), torch.zeros_like(rho_t))


def b3lyp_exhf_densities(rho, grad_rho, lapl, clip: float = CLIP):
    """popular_functionals.py:313-326 -- columns [lsda_x, b88_x, vwn_c, lyp_c]."""
    return torch.stack(
        (lsda_x_e(rho, clip), b88_x_e(rho, grad_rho, clip), vwn_c_e(rho, clip), lyp_c_e(rho, grad_rho, lapl, clip)),
        dim=1,
    )


def b3lyp_combine(features, ehf):
    """popular_functionals.py:335-338 -- append sum_{w,s} e_HF as the last column."""
    return torch.cat([features, ehf.sum(dim=(0, 1)).unsqueeze(1)], dim=1)


B3LYP_COEFFS = [1 - 0.2, 0.72, 1 - 0.81, 0.81, 0.2]  # popular_functionals.py:344-347


# --------------------------------------------------------------------------------------------
# DM21-style features  (grad_dft/functional.py:504-675, 797-822)
# --------------------------------------------------------------------------------------------
def dm21_coefficient_inputs(rho, grad_rho, tau, clip: float = CLIP):
    """functional.py:520-531 -- [rho_a, rho_b, |g_a+g_b|^2, |g_a|^2, |g_b|^2, tau_a, tau_b]."""
    rho = torch.maximum(rho.abs(), torch.full_like(rho, clip)) * torch.sign(rho)
    gnorm = (grad_rho ** 2).sum(dim=-1)
    gnorm_ss = (grad_rho.sum(dim=1, keepdim=True) ** 2).sum(dim=-1)
    return torch.cat((rho, gnorm_ss, gnorm, tau), dim=1)


def dm21_densities(rho, grad_rho, tau, functional_type: str = "LDA", clip: float = CLIP):
    """functional.py:573-626."""
    beta = 1 / 1024.0
    ranges = {"LDA": (1, 1), "DM21": (1, 1), "GGA": (2, 1), "MGGA": (2, 2)}
    nu, nw = ranges[functional_type]
    sigma = (grad_rho ** 2).sum(dim=-1)
    log_rho = torch.log2(torch.clamp(rho, min=clip))
    log_g = torch.log2(torch.clamp(sigma, min=clip)) / 2
    log_x = log_g - 4 / 3.0 * log_rho
    live = log_rho > math.log2(clip)
    zero = torch.zeros_like(log_rho)
    log_u = torch.where(live, log_x - torch.log2(1 + beta * torch.exp2(log_x)) + math.log2(beta), zero)
    log_tau = torch.log2(torch.clamp(tau, min=clip))
    log_1t = -(5 / 3.0 * log_rho - log_tau + 2 / 3.0 * math.log2(6 * math.pi ** 2) + math.log2(3 / 5.0))
    log_w = torch.where(live, log_1t - torch.log2(1 + beta * torch.exp2(log_1t)) + math.log2(beta), zero)
    cols = []
    for i in range(nu):
        for j in range(nw):
            term = torch.exp2(4 / 3.0 * log_rho + i * log_u + j * log_w).sum(dim=1, keepdim=True)
            if i == 0 and j == 0:
                term = term * (-2 * math.pi * (3 / (4 * math.pi)) ** (4 / 3))
            cols.append(term)
    return torch.cat(cols, dim=1)


def dm21_combine_cinputs(cinputs, ehf):
    """functional.py:649 -- HF features appended by spin: [w0 a, w1 a, w0 b, w1 b]."""
    return torch.cat([cinputs, ehf[:, 0].T, ehf[:, 1].T], dim=1)


def dm21_combine_densities(densities, ehf):
    """functional.py:673-675 -- one spin-summed HF column per omega."""
    return torch.cat([densities] + [ehf[i].sum(dim=0, keepdim=True).T for i in range(ehf.shape[0])], dim=1)


def dm21_mlp_init(n_in: int = 11, width: int = 256, n_layers: int = 6, n_out: int = 3, seed: int = 1984):
    """Seeded stand-in for DM21's weights: He-normal kernels, zero biases, identity added to the square
    residual kernels (functional.py:913-921), LayerNorm scale 1 / bias 0 (flax defaults)."""
    g = torch.Generator().manual_seed(seed)
    p = {}

    def he(i, o):
        return torch.randn(i, o, generator=g, dtype=F64) * math.sqrt(2.0 / i)

    p["Dense_0.kernel"], p["Dense_0.bias"] = he(n_in, width), torch.zeros(width, dtype=F64)
    for k in range(n_layers):
        p[f"Dense_{k + 1}.kernel"] = he(width, width) + torch.eye(width, dtype=F64)
        p[f"Dense_{k + 1}.bias"] = torch.zeros(width, dtype=F64)
        p[f"LayerNorm_{k}.scale"] = torch.ones(width, dtype=F64)
        p[f"LayerNorm_{k}.bias"] = torch.zeros(width, dtype=F64)
    p[f"Dense_{n_layers + 1}.kernel"] = he(width, n_out)
    p[f"Dense_{n_layers + 1}.bias"] = torch.zeros(n_out, dtype=F64)
    return p


def dm21_mlp(params, x, squash_offset: float = 1e-4, sigmoid_scale: float = 2.0):
    """functional.py:797-822 + head 407-419: log|x|+eps -> dense -> tanh -> 6x(dense+res -> LayerNorm
    (eps 1e-6, flax default) -> ELU) -> dense -> scaled sigmoid."""
    n_layers = sum(1 for k in params if k.startswith("LayerNorm_") and k.endswith(".scale"))
    x = torch.log(x.abs() + squash_offset)
    x = torch.tanh(x @ params["Dense_0.kernel"] + params["Dense_0.bias"])
    for k in range(n_layers):
        y = x @ params[f"Dense_{k + 1}.kernel"] + params[f"Dense_{k + 1}.bias"] + x
        mu = y.mean(dim=-1, keepdim=True)
        var = ((y - mu) ** 2).mean(dim=-1, keepdim=True)
        y = (y - mu) * torch.rsqrt(var + 1e-6) * params[f"LayerNorm_{k}.scale"] + params[f"LayerNorm_{k}.bias"]
        x = torch.nn.functional.elu(y)
    x = x @ params[f"Dense_{n_layers + 1}.kernel"] + params[f"Dense_{n_layers + 1}.bias"]
    return sigmoid_scale * torch.sigmoid(x / sigmoid_scale)


def mgga_feature_densities(rho, grad_rho, tau, functional_type: str = "MGGA", clip: float = CLIP):
    """functional.py:1087-1202 (`densities`, row f3): per-spin u/w expansion of rho^{4/3}; the
    correlation half is identically zero upstream (round(.,-30) -> 0, then a `> clip` test on it), which
    is reproduced here as explicit zeros with zero gradient."""
    beta = 1 / 1024.0
    ranges = {"LDA": (1, 1), "DM21": (1, 1), "GGA": (2, 1), "MGGA": (2, 2)}
    nu, nw = ranges[functional_type]
    sigma = (grad_rho ** 2).sum(dim=-1)
    log_rho = torch.log2(torch.clamp(rho, min=clip))
    log_g = torch.log2(torch.clamp(sigma, min=clip)) / 2
    log_x = log_g - 4 / 3.0 * log_rho
    live = log_rho > math.log2(clip)
    zero = torch.zeros_like(log_rho)
    log_u = torch.where(live, log_x - torch.log2(1 + beta * torch.exp2(log_x)) + math.log2(beta), zero)
    log_tau = torch.log2(torch.clamp(tau, min=clip))
    log_1t = log_tau - 5 / 3.0 * log_rho
    log_w = torch.where(live, log_1t - torch.log2(1 + beta * torch.exp2(log_1t)) + math.log2(beta), zero)
    cols = []
    for i in range(nu):
        for j in range(nw):
            cols.append(torch.exp2(4 / 3.0 * log_rho + i * log_u + j * log_w))
    for _ in range(nu * nw):
        cols.append(torch.zeros_like(rho))
    return torch.cat(cols, dim=1)


# --------------------------------------------------------------------------------------------
# closed-form VJP of the density family -- used to cross-check autograd
# --------------------------------------------------------------------------------------------
def density_vjp_formula(ao, grad_ao, lap_ao, rho_bar=None, grho_bar=None, tau_bar=None, lapl_bar=None):
    """Dbar_s = ao^T (rb_s*ao + 2 sum_j gb_sj*dj_ao + 2 lb_s*lap_ao) + sum_j dj_ao^T ((tb_s/2 + 2 lb_s)*dj_ao),
    with lap_ao[r,b] = sum_i grad_2_ao[r,b,i].  Un-symmetrised (index a on the left operand)."""
    N, n = ao.shape
    out = torch.zeros(2, n, n, dtype=ao.dtype)
    for s in range(2):
        M = torch.zeros_like(ao)
        if rho_bar is not None:
            M = M + rho_bar[:, s, None] * ao
        if grho_bar is not None:
            M = M + 2.0 * torch.einsum("rj,rbj->rb", grho_bar[:, s], grad_ao)
        if lapl_bar is not None:
            M = M + 2.0 * lapl_bar[:, s, None] * lap_ao
        out[s] = ao.T @ M
        k = None
        if tau_bar is not None:
            k = 0.5 * tau_bar[:, s]
        if lapl_bar is not None:
            k = 2.0 * lapl_bar[:, s] if k is None else k + 2.0 * lapl_bar[:, s]
        if k is not None:
            out[s] = out[s] + torch.einsum("raj,r,rbj->ab", grad_ao, k, grad_ao)
    return out


# --------------------------------------------------------------------------------------------
# predictor assembly  (grad_dft/train.py:86-216)
# --------------------------------------------------------------------------------------------
def xc_energy_of_rdm1(rdm1, mol: dict, functional: str, params=None, clip: float = CLIP):
    """train.py:114-121 -- compute_densities + compute_coefficient_inputs + xc_energy for a named
    functional.  ``mol`` is a dict of tensors with the reference's Molecule field names
    (molecule.py:76-102) plus 'weights'.  HF pieces enter under stop_gradient (functional.py:176,203)."""
    ao, gao, w = mol["ao"], mol["grad_ao"], mol["weights"]
    if functional in ("LSDA", "B88", "VWN", "LYP", "PW92"):
        rho = density(rdm1, ao)
        if functional == "LSDA":
            d = lsda_x_e(rho, clip).unsqueeze(1)
        elif functional == "B88":
            d = torch.stack((lsda_x_e(rho, clip), b88_x_e(rho, grad_density(rdm1, ao, gao), clip)), dim=1)
        elif functional == "VWN":
            d = vwn_c_e(rho, clip).unsqueeze(1)
        elif functional == "PW92":
            d = pw92_c_e(rho, clip).unsqueeze(1)
        else:
            d = lyp_c_e(rho, grad_density(rdm1, ao, gao), lapl_density(rdm1, ao, gao, mol["grad_n_ao2"]), clip).unsqueeze(1)
        d = abs_clip(d, clip)
        c = torch.ones(1, d.shape[1], dtype=d.dtype)
        if functional == "B88":
            c = torch.ones(1, 1, dtype=d.dtype)  # popular_functionals.py:376: coefficient [[1.0]] broadcast over 2 columns
            c = c.expand(1, 2)
        return xc_energy(c, d, w, clip)
    if functional == "B3LYP":
        rho = density(rdm1, ao)
        grho = grad_density(rdm1, ao, gao)
        lapl = lapl_density(rdm1, ao, gao, mol["grad_n_ao2"])
        feats = b3lyp_exhf_densities(rho, grho, lapl, clip)
        ehf = HF_energy_density(rdm1, ao, mol["chi"][:, :1]).detach()
        d = abs_clip(b3lyp_combine(feats, ehf), clip)
        c = torch.tensor([B3LYP_COEFFS], dtype=d.dtype)
        return xc_energy(c, d, w, clip)
    if functional == "DM21":
        rho = density(rdm1, ao)
        grho = grad_density(rdm1, ao, gao)
        tau = kinetic_density(rdm1, gao)
        ehf = HF_energy_density(rdm1, ao, mol["chi"]).detach()
        d = abs_clip(dm21_combine_densities(dm21_densities(rho, grho, tau, "LDA", clip), ehf), clip)
        ci = dm21_combine_cinputs(dm21_coefficient_inputs(rho, grho, tau, clip), ehf)
        return xc_energy(dm21_mlp(params, ci), d, w, clip)
    raise ValueError(functional)


def _fock_common(rdm1, mol, fock_xc, clip):
    """train.py:148-163 -- h1e + J + Dbar, aclip, symmetrise, aclip."""
    P = rdm1.sum(dim=0)
    fock = mol["h1e"] + coulomb_potential(P, mol["rep_tensor"]) + fock_xc
    fock = abs_clip(fock, clip)
    fock = 0.5 * (fock + fock.transpose(1, 2))
    return abs_clip(fock, clip)


def predict_semilocal(mol: dict, functional: str, clip: float = CLIP):
    """train.py:147-163,215-216 for functionals without explicit HF terms."""
    rdm1 = mol["rdm1"].detach().clone().requires_grad_(True)
    Exc = xc_energy_of_rdm1(rdm1, mol, functional, clip=clip)
    (fock_xc,) = torch.autograd.grad(Exc, rdm1)
    P = mol["rdm1"].sum(dim=0)
    energy = Exc.detach() + nonXC(P, mol["h1e"], mol["rep_tensor"], mol["nuclear_repulsion"])
    return energy, abs_clip(_fock_common(mol["rdm1"], mol, fock_xc, clip), clip)


def predict_b3lyp(mol: dict, clip: float = CLIP):
    """train.py:147-216 for B3LYP: autodiff part + explicit HF Fock term
    (popular_functionals.py:357-372 -> molecule.py:600-613), F += V + V^T, aclip."""
    rdm1 = mol["rdm1"].detach().clone().requires_grad_(True)
    Exc = xc_energy_of_rdm1(rdm1, mol, "B3LYP", clip=clip)
    (fock_xc,) = torch.autograd.grad(Exc, rdm1)
    D = mol["rdm1"]
    P = D.sum(dim=0)
    energy = Exc.detach() + nonXC(P, mol["h1e"], mol["rep_tensor"], mol["nuclear_repulsion"])
    fock = _fock_common(D, mol, fock_xc, clip)
    ao, gao = mol["ao"], mol["grad_ao"]
    feats = b3lyp_exhf_densities(density(D, ao), grad_density(D, ao, gao), lapl_density(D, ao, gao, mol["grad_n_ao2"]), clip)
    chi = mol["chi"][:, :1]
    ehf = HF_energy_density(D, ao, chi).detach().requires_grad_(True)
    c = torch.tensor([B3LYP_COEFFS], dtype=ao.dtype)
    E = xc_energy(c, b3lyp_combine(feats, ehf), mol["weights"], clip)  # molecule.py:600-604 (no aclip on densities here)
    (g,) = torch.autograd.grad(E, ehf)
    v = HF_fock(chi, g, ao).sum(dim=0)
    fock = abs_clip(fock + (v + v.transpose(1, 2)), clip)  # train.py:205,212: fock += V + V^T
    return energy, abs_clip(fock, clip)


def predict_dm21(mol: dict, params, clip: float = CLIP):
    """train.py:147-216 for a DM21-shaped functional: both explicit HF terms
    (functional.py:714-717, 755-758)."""
    rdm1 = mol["rdm1"].detach().clone().requires_grad_(True)
    Exc = xc_energy_of_rdm1(rdm1, mol, "DM21", params=params, clip=clip)
    (fock_xc,) = torch.autograd.grad(Exc, rdm1)
    D = mol["rdm1"]
    P = D.sum(dim=0)
    energy = Exc.detach() + nonXC(P, mol["h1e"], mol["rep_tensor"], mol["nuclear_repulsion"])
    fock = _fock_common(D, mol, fock_xc, clip)
    ao, gao, chi, w = mol["ao"], mol["grad_ao"], mol["chi"], mol["weights"]
    rho, grho, tau = density(D, ao), grad_density(D, ao, gao), kinetic_density(D, gao)
    grad_densities = dm21_densities(rho, grho, tau, "LDA", clip)
    grad_cinputs = dm21_coefficient_inputs(rho, grho, tau, clip)
    ehf0 = HF_energy_density(D, ao, chi).detach()
    densities = dm21_combine_densities(grad_densities, ehf0)
    cinputs = dm21_combine_cinputs(grad_cinputs, ehf0)
    # densitygrads: d E / d ehf through the densities only (molecule.py:600-604)
    ehf = ehf0.clone().requires_grad_(True)
    E = xc_energy(dm21_mlp(params, cinputs), dm21_combine_densities(grad_densities, ehf), w, clip)
    (g,) = torch.autograd.grad(E, ehf)
    v = HF_fock(chi, g, ao).sum(dim=0)
    fock = abs_clip(fock + (v + v.transpose(1, 2)), clip)  # train.py:205,212: fock += V + V^T
    # coefficient_input_grads: d E / d ehf through the network inputs only (molecule.py:672-676)
    ehf = ehf0.clone().requires_grad_(True)
    E = xc_energy(dm21_mlp(params, dm21_combine_cinputs(grad_cinputs, ehf)), densities, w, clip)
    (g,) = torch.autograd.grad(E, ehf)
    v = HF_fock(chi, g, ao).sum(dim=0)
    fock = abs_clip(fock + (v + v.transpose(1, 2)), clip)  # train.py:205,212: fock += V + V^T
    return energy, abs_clip(fock, clip)


def predict_dm21_traced(mol: dict, params, clip: float = CLIP):
    """train.py:147-216 for a DM21-shaped functional as `jax.grad` of an enclosing function sees it: every
    intermediate keeps its dependence on (params, mol["rdm1"]) -- V_xc is itself differentiable (create_graph), the
    features passed to the explicit HF routes are NOT stopped (train.py:200-213) -- except what the reference wraps
    in stop_gradient (the HF energy density entering E_xc, functional.py:176,203, and the `ehf` argument of the two
    HF routes).  Used to check gradients through the SCF loops (evaluate.py:257-352, 917-1038)."""
    D = mol["rdm1"] if mol["rdm1"].requires_grad else mol["rdm1"].detach().clone().requires_grad_(True)
    Exc = xc_energy_of_rdm1(D, mol, "DM21", params=params, clip=clip)
    (fock_xc,) = torch.autograd.grad(Exc, D, create_graph=True)
    P = D.sum(dim=0)
    energy = Exc + nonXC(P, mol["h1e"], mol["rep_tensor"], mol["nuclear_repulsion"])
    fock = _fock_common(D, mol, fock_xc, clip)
    ao, gao, chi, w = mol["ao"], mol["grad_ao"], mol["chi"], mol["weights"]
    rho, grho, tau = density(D, ao), grad_density(D, ao, gao), kinetic_density(D, gao)
    grad_densities = dm21_densities(rho, grho, tau, "LDA", clip)
    grad_cinputs = dm21_coefficient_inputs(rho, grho, tau, clip)
    ehf0 = HF_energy_density(D, ao, chi).detach()
    densities = dm21_combine_densities(grad_densities, ehf0)
    cinputs = dm21_combine_cinputs(grad_cinputs, ehf0)
    ehf = ehf0.clone().requires_grad_(True)
    E = xc_energy(dm21_mlp(params, cinputs), dm21_combine_densities(grad_densities, ehf), w, clip)
    (g,) = torch.autograd.grad(E, ehf, create_graph=True)
    v = HF_fock(chi, g, ao).sum(dim=0)
    fock = abs_clip(fock + (v + v.transpose(1, 2)), clip)
    ehf = ehf0.clone().requires_grad_(True)
    E = xc_energy(dm21_mlp(params, dm21_combine_cinputs(grad_cinputs, ehf)), densities, w, clip)
    (g,) = torch.autograd.grad(E, ehf, create_graph=True)
    v = HF_fock(chi, g, ao).sum(dim=0)
    fock = abs_clip(fock + (v + v.transpose(1, 2)), clip)
    return energy, abs_clip(fock, clip)


def diff_simple_scf_loop_energy(mol: dict, predict: Callable, cycles: int, mixing_factor: float = 0.4):
    """evaluate.py:300-350 -- linear density mixing; differentiable when `predict` is traced."""
    mol = dict(mol)
    e, fock = predict(mol)
    nelecs = mol["mo_occ"].sum(dim=1).round().to(torch.int64)
    for _ in range(cycles):
        mo_energy, mo_coeff = safe_fock_solver(fock, mol["s1e"])
        occ = get_occ(mo_energy.detach(), nelecs)
        mol["rdm1"] = (1 - mixing_factor) * mol["rdm1"] + mixing_factor * make_rdm1(mo_coeff, occ)
        e, fock = predict(mol)
    return e, mol


# --------------------------------------------------------------------------------------------
# SCF harness pieces (row f1)  (grad_dft/utils/eigenproblem.py, grad_dft/evaluate.py)
# --------------------------------------------------------------------------------------------
def safe_fock_solver(fock, overlap):
    """utils/eigenproblem.py:110-149 -- Cholesky-reduced generalised symmetric eigenproblem per spin
    (forward values only)."""
    L = torch.linalg.cholesky(overlap)
    Linv = torch.linalg.inv(L)
    es, cs = [], []
    for s in range(2):
        C = Linv @ fock[s] @ Linv.T
        e, v = torch.linalg.eigh(C)
        es.append(e)
        cs.append(Linv.T @ v)
    return torch.stack(es), torch.stack(cs)


def jittable_diis_run(overlap, rdm1, fock, energy, data, cycle: int, max_diis: int = 10):
    """evaluate.py:1111-1205 with A = identity (evaluate.py:972-973).  ``data`` =
    (density_vector, fock_vector, energy_vector, error_vector).  The `.at[cycle]` write with
    cycle == max_diis is out of bounds and dropped, as under JAX scatter semantics."""
    dv, fv, ev, errv = [t.clone() for t in data]
    fds = torch.einsum("sjk,skl,lm->sjm", fock, rdm1, overlap)
    err = fds - fds.transpose(1, 2)
    if cycle > max_diis:
        errv = torch.cat([errv, err[None]])[1:]
        dv = torch.cat([dv, rdm1[None]])[1:]
        fv = torch.cat([fv, fock[None]])[1:]
        ev = torch.cat([ev, energy.reshape(1)])[1:]
    elif cycle < max_diis:
        errv[cycle], dv[cycle], fv[cycle], ev[cycle] = err, rdm1, fock, energy
    m = errv.shape[0]
    G = torch.einsum("iskl,jskl->sij", errv, errv)
    B = torch.zeros(2, m + 1, m + 1, dtype=fock.dtype)
    B[:, 1:, 1:] = G
    for i in range(m + 2):  # evaluate.py:1195: loop runs to m+2; out-of-range writes are dropped
        val = 1.0 if i <= cycle else 0.0
        if i + 1 <= m:
            B[:, 0, i + 1] = val
            B[:, i + 1, 0] = val
    for i in range(m + 2):
        if i + 1 <= m:
            B[:, i + 1, i + 1] = G[:, i, i] if i <= cycle else torch.ones(2, dtype=fock.dtype)
    Cv = torch.zeros(m + 1, dtype=fock.dtype)
    Cv[0] = 1.0
    x = torch.stack([torch.linalg.inv(B[s]) @ Cv for s in range(2)])[:, 1:]
    F = torch.einsum("si,isjk->sjk", x, fv)
    return F, (dv, fv, ev, errv)


def diff_scf_loop_energy(mol: dict, predict: Callable, cycles: int, max_diis: int = 10):
    """evaluate.py:965-1035 -- DIIS SCF; returns the energy after ``cycles`` iterations (the extra final
    iteration is computed and discarded upstream, evaluate.py:1030-1031) and the final dict."""
    mol = dict(mol)
    n = mol["s1e"].shape[0]
    e, fock = predict(mol)
    mol["fock"] = fock
    z = torch.zeros(max_diis, 2, n, n, dtype=fock.dtype)
    data = (z.clone(), z.clone(), torch.zeros(max_diis, dtype=fock.dtype), z.clone())
    nelecs = mol["mo_occ"].sum(dim=1).round().to(torch.int64)
    for cycle in range(cycles):
        F, data = jittable_diis_run(mol["s1e"], mol["rdm1"], mol["fock"], e, data, cycle, max_diis)
        mo_energy, mo_coeff = safe_fock_solver(F, mol["s1e"])
        mol["mo_energy"], mol["mo_coeff"] = mo_energy, mo_coeff
        mol["mo_occ"] = get_occ(mo_energy, nelecs)
        mol["rdm1"] = make_rdm1(mo_coeff, mol["mo_occ"])
        e, fock = predict(mol)
        mol["fock"] = fock
    return e, mol


# --------------------------------------------------------------------------------------------
# chi generation tail  (grad_dft/interface/pyscf.py)
# --------------------------------------------------------------------------------------------
def generate_chi_tensor(rdm1: torch.Tensor, ao: torch.Tensor, grid_coords: torch.Tensor, nu_fn: Callable, omegas: Sequence[float],
                        chunk_size: Optional[int] = 1024) -> torch.Tensor:
    """interface/pyscf.py:1110-1124 -- chi[r, w, s, a] = einsum("sbd,b,da->sa", rdm1, ao[r], nu_w[r]) chunk by chunk
    (`_nu_chunk`, external/_hf_density.py:69-103, yields nu for `chunk_size` grid points at a time; nu_fn(coords, omega)
    stands for libcint's int1e_grids), concatenated over chunks and stacked over omegas on axis 1."""
    chi = []
    N = ao.shape[0]
    chunk_size = chunk_size or N
    for omega in omegas:
        if omega < 0:
            raise ValueError("Range-separated parameter omega must be non-negative!")  # _hf_density.py:92-93
        parts = []
        for i in range(0, N, chunk_size):
            j = min(i + chunk_size, N)
            nu = torch.as_tensor(nu_fn(grid_coords[i:j], omega), dtype=F64)
            parts.append(torch.einsum("sbd,rb,rda->rsa", rdm1, ao[i:j], nu))
        chi.append(torch.cat(parts, dim=0))
    return torch.stack(chi, dim=1) if chi else torch.zeros(0, dtype=F64)
