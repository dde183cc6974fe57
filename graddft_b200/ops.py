"""Differentiable ops over the C-ABI kernels (torch.autograd.Function bindings; no arithmetic of their own
beyond bookkeeping).

Every linear kernel is bound together with its transpose so that each is the other's VJP
(`density_forward` <-> `density_transpose`, `coulomb_j` <-> its transpose, `hf_fock` <-> the HF
energy-density contraction): autograd of any order closes over the same two kernels, which is what
training through the SCF loop needs (grad_dft/evaluate.py:917-1038).
"""
from __future__ import annotations

import contextlib

from typing import Sequence, Optional, Tuple

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
from ._lib import GDFT_GRAD, GDFT_HF, GDFT_LAPL, GDFT_RHO, GDFT_TAU, check, lib, ptr, stream_ptr, workspace, wptr

F64 = torch.float64

# Optional timing hook (bench.py): when a dict is installed, the named C-ABI calls are bracketed by CUDA events
# on the launching stream and the (start, end) pairs appended to TIMING[name].  None = no overhead.
TIMING: Optional[dict] = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if TIMING is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if TIMING is not None:
            self.b.record()
            TIMING.setdefault(self.name, []).append((self.a, self.b))
        return False


def _c(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dtype != F64:
        raise TypeError(f"graddft_b200 computes in float64, got {t.dtype}")
    return t.contiguous()


class PackedBasis:
    """Planar, column-padded copy of the per-molecule constants ao / grad_ao / grad_n_ao[2] (/ chi).

    Built once per molecule (they do not change across SCF iterations or training steps) by
    gdft_pack_basis / gdft_pack_chi: planes[C, N, npad] with plane 0 = ao, 1..3 = d/dx,y,z ao,
    4 = sum_i grad_n_ao[2][..., i]; chi_packed[W, 2, N, npad].
    """

    def __init__(self, ao: torch.Tensor, grad_ao: Optional[torch.Tensor] = None, grad2_ao: Optional[torch.Tensor] = None,
                 chi: Optional[torch.Tensor] = None):
        L = lib()
        ao = _c(ao)
        if not ao.is_cuda:
            raise _lib.GdftError("graddft_b200 kernels need CUDA tensors (there is no CPU path)")
        self.N, self.n = int(ao.shape[0]), int(ao.shape[1])
        self.device = ao.device
        self.npad = int(L.gdft_npad(self.n))
        self.nplanes = 1 if grad_ao is None else (4 if grad2_ao is None else 5)
        if grad2_ao is not None and grad_ao is None:
            raise ValueError("grad2_ao needs grad_ao")
        self.planes = torch.empty((self.nplanes, self.N, self.npad), dtype=F64, device=self.device)
        grad_ao, grad2_ao = _c(grad_ao), _c(grad2_ao)
        for t, shape in ((grad_ao, (self.N, self.n, 3)), (grad2_ao, (self.N, self.n, 3))):
            if t is not None and tuple(t.shape) != shape:
                raise TypeError(f"expected shape {shape}, got {tuple(t.shape)}")
        check(L.gdft_pack_basis(stream_ptr(), self.N, self.n, ptr(ao), ptr(grad_ao), ptr(grad2_ao), ptr(self.planes), self.nplanes),
              "gdft_pack_basis")
        self.W = 0
        self.chi_packed = None
        if chi is not None:
            self.set_chi(chi)

    def set_chi(self, chi: torch.Tensor) -> None:
        chi = _c(chi)
        if chi.dim() != 4 or chi.shape[0] != self.N or chi.shape[2] != 2 or chi.shape[3] != self.n:
            raise TypeError(f"chi must be [N, omega, 2, n], got {tuple(chi.shape)}")
        self.W = int(chi.shape[1])
        self.chi_packed = torch.empty((self.W, 2, self.N, self.npad), dtype=F64, device=self.device)
        check(lib().gdft_pack_chi(stream_ptr(), self.N, self.n, self.W, ptr(chi), ptr(self.chi_packed)), "gdft_pack_chi")

    def select_chi(self, indices) -> "PackedBasis":
        """A view of this basis whose chi planes are the given omega indices (Molecule.select_HF_omegas).
        Contiguous ascending index ranges are zero-copy slices; other selections are gathered once and kept."""
        idx = tuple(int(i) for i in indices)
        if idx == tuple(range(self.W)):
            return self
        cache = self.__dict__.setdefault("_selected", {})
        if idx not in cache:
            other = object.__new__(PackedBasis)
            other.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_selected"})
            if idx == tuple(range(idx[0], idx[0] + len(idx))):
                other.chi_packed = self.chi_packed[idx[0]:idx[0] + len(idx)]
            else:
                other.chi_packed = self.chi_packed[list(idx)].contiguous()
            other.W = len(idx)
            cache[idx] = other
        return cache[idx]


# ---------------------------------------------------------------------------------------------------------
# density family: L (rdm1 -> grid) and L^T (grid cotangents -> rdm1 cotangent)
# ---------------------------------------------------------------------------------------------------------
def _density_fwd_raw(basis: PackedBasis, rdm1: torch.Tensor, flags: int):
    L = lib()
    rdm1 = _c(rdm1)
    if tuple(rdm1.shape) != (2, basis.n, basis.n):
        raise TypeError(f"rdm1 must be [2, {basis.n}, {basis.n}], got {tuple(rdm1.shape)}")
    N, dev = basis.N, basis.device
    rho = torch.empty((N, 2), dtype=F64, device=dev) if flags & GDFT_RHO else None
    grho = torch.empty((N, 2, 3), dtype=F64, device=dev) if flags & GDFT_GRAD else None
    tau = torch.empty((N, 2), dtype=F64, device=dev) if flags & GDFT_TAU else None
    lapl = torch.empty((N, 2), dtype=F64, device=dev) if flags & GDFT_LAPL else None
    ehf = torch.empty((basis.W, 2, N), dtype=F64, device=dev) if flags & GDFT_HF else None
    if (flags & GDFT_HF) and basis.chi_packed is None:
        raise ValueError("Precomputed chi tensor has not been loaded.")
    ws = workspace(L.gdft_workspace_bytes(_lib.OP_DENSITY_FWD, N, basis.n, flags, basis.W), dev)
    with _timed("gdft_density_fwd"):
        check(L.gdft_density_fwd(stream_ptr(), N, basis.n, flags, basis.nplanes, ptr(basis.planes), ptr(rdm1),
                                 ptr(basis.chi_packed) if flags & GDFT_HF else None, basis.W,
                                 ptr(rho), ptr(grho), ptr(tau), ptr(lapl), ptr(ehf), wptr(ws), ws.numel()), "gdft_density_fwd")
    return rho, grho, tau, lapl, ehf


_BWD_OUT: Optional[torch.Tensor] = None


@contextlib.contextmanager
def density_bwd_into(view: Optional[torch.Tensor]):
    """The next density VJP (gdft_density_bwd) inside this context writes its [2, n, n] result straight into `view`
    (2*n*n contiguous float64 elements of device memory, e.g. a segment of the all-reduce payload of
    `distributed.FockComm`) instead of a fresh tensor: the split-K second-stage reduce becomes the producer of the exchange
    buffer.  Consumed by the first matching call only; callers check `result.data_ptr()` to see whether it was used."""
    global _BWD_OUT
    prev, _BWD_OUT = _BWD_OUT, view
    try:
        yield
    finally:
        _BWD_OUT = prev


def _density_bwd_raw(basis: PackedBasis, flags: int, rho_bar, grho_bar, tau_bar, lapl_bar) -> torch.Tensor:
    global _BWD_OUT
    L = lib()
    N, dev = basis.N, basis.device
    out = None
    if _BWD_OUT is not None and _BWD_OUT.numel() == 2 * basis.n * basis.n and _BWD_OUT.device == dev and _BWD_OUT.is_contiguous():
        out, _BWD_OUT = _BWD_OUT.view(2, basis.n, basis.n), None
    if out is None:
        out = torch.empty((2, basis.n, basis.n), dtype=F64, device=dev)
    ws = workspace(L.gdft_workspace_bytes(_lib.OP_DENSITY_BWD, N, basis.n, flags, 0), dev)
    rho_bar, grho_bar, tau_bar, lapl_bar = _c(rho_bar), _c(grho_bar), _c(tau_bar), _c(lapl_bar)
    with _timed("gdft_density_bwd"):
        check(L.gdft_density_bwd(stream_ptr(), N, basis.n, flags, basis.nplanes, ptr(basis.planes), ptr(rho_bar), ptr(grho_bar),
                                 ptr(tau_bar), ptr(lapl_bar), ptr(out), wptr(ws), ws.numel()), "gdft_density_bwd")
    return out


def _hf_fock_raw(basis: PackedBasis, g: torch.Tensor) -> torch.Tensor:
    L = lib()
    g = _c(g)
    if tuple(g.shape) != (basis.W, 2, basis.N):
        raise TypeError(f"g must be [{basis.W}, 2, {basis.N}], got {tuple(g.shape)}")
    out = torch.empty((basis.W, 2, basis.n, basis.n), dtype=F64, device=basis.device)
    ws = workspace(L.gdft_workspace_bytes(_lib.OP_HF_FOCK, basis.N, basis.n, 0, basis.W), basis.device)
    with _timed("gdft_hf_fock"):
        check(L.gdft_hf_fock(stream_ptr(), basis.N, basis.n, basis.W, basis.nplanes, ptr(basis.planes), ptr(basis.chi_packed), ptr(g),
                             ptr(out), wptr(ws), ws.numel()), "gdft_hf_fock")
    return out


class _DensityForward(Function):
    @staticmethod
    def forward(ctx, rdm1, basis, flags):
        ctx.basis, ctx.flags = basis, flags
        ctx.set_materialize_grads(False)  # unused outputs (e.g. the stop_gradient'ed e_HF) must not cost a GEMM in the VJP
        return _density_fwd_raw(basis, rdm1, flags)

    @staticmethod
    def backward(ctx, rho_bar, grho_bar, tau_bar, lapl_bar, ehf_bar):
        basis = ctx.basis
        flags = 0
        if rho_bar is not None: flags |= GDFT_RHO
        if grho_bar is not None: flags |= GDFT_GRAD
        if tau_bar is not None: flags |= GDFT_TAU
        if lapl_bar is not None: flags |= GDFT_LAPL
        dbar = None
        if flags:
            dbar = density_transpose(basis, rho_bar, grho_bar, tau_bar, lapl_bar)
        if ehf_bar is not None:
            # e_HF[w,s,r] = -1/2 rowdot(chi_ws, ao D_s): its transpose is the HF Fock contraction summed over omega
            v = hf_fock(basis, ehf_bar).sum(dim=0)
            dbar = v if dbar is None else dbar + v
        return dbar, None, None


class _DensityTranspose(Function):
    @staticmethod
    def forward(ctx, basis, flags, rho_bar, grho_bar, tau_bar, lapl_bar):
        ctx.basis, ctx.flags = basis, flags
        return _density_bwd_raw(basis, flags, rho_bar, grho_bar, tau_bar, lapl_bar)

    @staticmethod
    def backward(ctx, dd):
        rho, grho, tau, lapl, _ = _DensityForward.apply(dd, ctx.basis, ctx.flags)
        return None, None, rho, grho, tau, lapl


class _HFFock(Function):
    @staticmethod
    def forward(ctx, g, basis):
        ctx.basis = basis
        return _hf_fock_raw(basis, g)

    @staticmethod
    def backward(ctx, fbar):
        basis = ctx.basis
        rows = []
        for w in range(basis.W):
            ehf = _DensityForward.apply(fbar[w], basis, GDFT_HF)[4]
            rows.append(ehf[w])
        return torch.stack(rows, dim=0), None


def density_forward(basis: PackedBasis, rdm1: torch.Tensor, flags: int):
    """(rho, grad_rho, tau, lapl, ehf) for the quantities selected by `flags` (others None)."""
    return _DensityForward.apply(rdm1, basis, flags)


def density_transpose(basis: PackedBasis, rho_bar=None, grho_bar=None, tau_bar=None, lapl_bar=None) -> torch.Tensor:
    flags = 0
    if rho_bar is not None: flags |= GDFT_RHO
    if grho_bar is not None: flags |= GDFT_GRAD
    if tau_bar is not None: flags |= GDFT_TAU
    if lapl_bar is not None: flags |= GDFT_LAPL
    if not flags:
        raise ValueError("density_transpose needs at least one cotangent")
    return _DensityTranspose.apply(basis, flags, rho_bar, grho_bar, tau_bar, lapl_bar)


def hf_fock(basis: PackedBasis, g: torch.Tensor) -> torch.Tensor:
    """F[w,s,a,c] = -1/2 sum_r ao[r,a] g[w,s,r] chi[r,w,s,c]."""
    if basis.chi_packed is None:
        raise ValueError("Precomputed chi tensor has not been loaded.")
    return _HFFock.apply(g, basis)


def hf_fock_sum(basis: PackedBasis, g: torch.Tensor) -> torch.Tensor:
    """sum_w F[w] of `hf_fock` -- what the hybrid wiring does with it (grad_dft/functional.py:714-717, 755-758) -- with the sum
    over omega taken inside the GEMM for up to two omegas (gdft_hf_fock_sum); no autograd (first-order predictor path)."""
    if basis.chi_packed is None:
        raise ValueError("Precomputed chi tensor has not been loaded.")
    g = _c(g.detach())
    if tuple(g.shape) != (basis.W, 2, basis.N):
        raise TypeError(f"g must be [{basis.W}, 2, {basis.N}], got {tuple(g.shape)}")
    if basis.W > 2:
        return _hf_fock_raw(basis, g).sum(dim=0)
    out = torch.empty((2, basis.n, basis.n), dtype=F64, device=basis.device)
    ws = workspace(lib().gdft_workspace_bytes(_lib.OP_HF_FOCK, basis.N, basis.n, 0, basis.W), basis.device)
    with _timed("gdft_hf_fock"):
        check(lib().gdft_hf_fock_sum(stream_ptr(), basis.N, basis.n, basis.W, basis.nplanes, ptr(basis.planes), ptr(basis.chi_packed), ptr(g),
                                     ptr(out), wptr(ws), ws.numel()), "gdft_hf_fock_sum")
    return out


# ---------------------------------------------------------------------------------------------------------
# ERI sweep
# ---------------------------------------------------------------------------------------------------------
def _eri_j_raw(P, eri, want_energy=False):
    L = lib()
    P, eri = _c(P), _c(eri)
    n = int(P.shape[0])
    if tuple(eri.shape) != (n, n, n, n) or tuple(P.shape) != (n, n):
        raise TypeError(f"rep_tensor must be [n,n,n,n] and rdm1 [n,n]; got {tuple(eri.shape)}, {tuple(P.shape)}")
    J = torch.empty((n, n), dtype=F64, device=P.device)
    EJ = torch.empty((1,), dtype=F64, device=P.device) if want_energy else None
    with _timed("gdft_eri_jk"):
        check(L.gdft_eri_jk(stream_ptr(), n, ptr(eri), ptr(P), ptr(J), None, ptr(EJ), None, 0), "gdft_eri_jk")
    return J, EJ


def _eri_jt_raw(Jbar, eri):
    L = lib()
    Jbar, eri = _c(Jbar), _c(eri)
    n = int(Jbar.shape[0])
    out = torch.empty((n, n), dtype=F64, device=Jbar.device)
    ws = workspace(L.gdft_workspace_bytes(_lib.OP_ERI_J, 0, n, 0, 0), Jbar.device)
    check(L.gdft_eri_j_transpose(stream_ptr(), n, ptr(eri), ptr(Jbar), ptr(out), wptr(ws), ws.numel()), "gdft_eri_j_transpose")
    return out


class _CoulombJ(Function):
    @staticmethod
    def forward(ctx, P, eri):
        ctx.save_for_backward(eri)
        return _eri_j_raw(P, eri)[0]

    @staticmethod
    def backward(ctx, Jbar):
        (eri,) = ctx.saved_tensors
        return _CoulombJT.apply(Jbar, eri), None


class _CoulombJT(Function):
    @staticmethod
    def forward(ctx, Jbar, eri):
        ctx.save_for_backward(eri)
        return _eri_jt_raw(Jbar, eri)

    @staticmethod
    def backward(ctx, g):
        (eri,) = ctx.saved_tensors
        return _CoulombJ.apply(g, eri), None


def coulomb_j(P: torch.Tensor, eri: torch.Tensor) -> torch.Tensor:
    return _CoulombJ.apply(P, eri)


def _eri_rows_shape(P, eri_rows):
    n = int(P.shape[0])
    if tuple(P.shape) != (n, n) or eri_rows.dim() != 3 or tuple(eri_rows.shape[1:]) != (n, n) or eri_rows.shape[0] > n * n:
        raise TypeError(f"rdm1 must be [n,n] and the rep_tensor row block [rows<=n*n, n, n]; got {tuple(P.shape)}, {tuple(eri_rows.shape)}")
    return n, int(eri_rows.shape[0])


class _CoulombJRows(Function):
    """J entries of a contiguous block of (p,q) rows of rep_tensor (row-sharded ERI)."""

    @staticmethod
    def forward(ctx, P, eri_rows):
        P, eri_rows = _c(P), _c(eri_rows)
        n, rows = _eri_rows_shape(P, eri_rows)
        ctx.save_for_backward(eri_rows)
        out = torch.empty((rows,), dtype=F64, device=P.device)
        with _timed("gdft_eri_jk"):
            check(lib().gdft_eri_j_rows(stream_ptr(), n, rows, ptr(eri_rows), ptr(P), ptr(out)), "gdft_eri_j_rows")
        return out

    @staticmethod
    def backward(ctx, Jbar_rows):
        (eri_rows,) = ctx.saved_tensors
        return _CoulombJRowsT.apply(Jbar_rows, eri_rows), None


class _CoulombJRowsT(Function):
    @staticmethod
    def forward(ctx, Jbar_rows, eri_rows):
        Jbar_rows = _c(Jbar_rows)
        n, rows = int(eri_rows.shape[1]), int(eri_rows.shape[0])
        ctx.save_for_backward(eri_rows)
        out = torch.empty((n, n), dtype=F64, device=eri_rows.device)
        ws = workspace(lib().gdft_workspace_bytes(_lib.OP_ERI_J, 0, n, 0, 0), eri_rows.device)
        check(lib().gdft_eri_j_transpose_rows(stream_ptr(), n, rows, ptr(eri_rows), ptr(Jbar_rows), ptr(out), wptr(ws), ws.numel()),
              "gdft_eri_j_transpose_rows")
        return out

    @staticmethod
    def backward(ctx, g):
        (eri_rows,) = ctx.saved_tensors
        return _CoulombJRows.apply(g, eri_rows), None


def coulomb_j_rows(P: torch.Tensor, eri_rows: torch.Tensor) -> torch.Tensor:
    """J[row0:row0+rows] (flattened (p,q) order) from the rank-local row block eri_rows[rows, n, n]."""
    return _CoulombJRows.apply(P, eri_rows)


def _eri_shapes(P, eri):
    n = int(P.shape[0])
    if tuple(eri.shape) != (n, n, n, n) or tuple(P.shape) != (n, n):
        raise TypeError(f"rep_tensor must be [n,n,n,n] and the matrix [n,n]; got {tuple(eri.shape)}, {tuple(P.shape)}")
    return n


def _eri_jk_raw(P, eri):
    """(J, K) from ONE pass over the tensor (gdft_eri_jk with K requested)."""
    L = lib()
    P, eri = _c(P), _c(eri)
    n = _eri_shapes(P, eri)
    J = torch.empty((n, n), dtype=F64, device=P.device)
    K = torch.empty((n, n), dtype=F64, device=P.device)
    ws = workspace(L.gdft_workspace_bytes(_lib.OP_ERI_J, 0, n, 0, 0), P.device)
    with _timed("gdft_eri_jk"):
        check(L.gdft_eri_jk(stream_ptr(), n, ptr(eri), ptr(P), ptr(J), ptr(K), None, wptr(ws), ws.numel()), "gdft_eri_jk")
    return J, K


def _eri_kt_raw(Kbar, eri):
    L = lib()
    Kbar, eri = _c(Kbar), _c(eri)
    n = _eri_shapes(Kbar, eri)
    out = torch.empty((n, n), dtype=F64, device=Kbar.device)
    ws = workspace(L.gdft_workspace_bytes(_lib.OP_ERI_J, 0, n, 0, 0), Kbar.device)
    check(L.gdft_eri_k_transpose(stream_ptr(), n, ptr(eri), ptr(Kbar), ptr(out), wptr(ws), ws.numel()), "gdft_eri_k_transpose")
    return out


class _CoulombJK(Function):
    """(J, K) of one sweep; each output's cotangent goes back through its own transposed sweep (both linear in P, so
    autograd of any order closes over the same four kernels)."""

    @staticmethod
    def forward(ctx, P, eri):
        ctx.save_for_backward(eri)
        ctx.set_materialize_grads(False)
        return _eri_jk_raw(P, eri)

    @staticmethod
    def backward(ctx, Jbar, Kbar):
        (eri,) = ctx.saved_tensors
        out = None
        if Jbar is not None:
            out = _CoulombJT.apply(Jbar, eri)
        if Kbar is not None:
            kt = _CoulombKT.apply(Kbar, eri)
            out = kt if out is None else out + kt
        return out, None


class _CoulombKT(Function):
    @staticmethod
    def forward(ctx, Kbar, eri):
        ctx.save_for_backward(eri)
        return _eri_kt_raw(Kbar, eri)

    @staticmethod
    def backward(ctx, g):
        (eri,) = ctx.saved_tensors
        return _CoulombJK.apply(g, eri)[1], None


def coulomb_jk(P: torch.Tensor, eri: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """J[p,q] = sum_rt (pq|rt) P[r,t] and K[p,r] = sum_qt (pq|rt) P[q,t] from one pass over rep_tensor (8 n^4 bytes for both);
    differentiable to any order w.r.t. P.  K is the exchange pairing named by BASELINE.json's north_star; the reference itself
    never contracts rep_tensor this way."""
    return _CoulombJK.apply(P, eri)


def coulomb_k(P: torch.Tensor, eri: torch.Tensor) -> torch.Tensor:
    """K[p,r] = sum_qt (pq|rt) P[q,t] (see coulomb_jk)."""
    return _CoulombJK.apply(P, eri)[1]


def coulomb_j_and_energy(P: torch.Tensor, eri: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """J and E_J = 1/2 <P,J> from one sweep (non-differentiable fast path used by the predictor)."""
    pe = packed_eri_for(eri)
    if pe is not None:
        return pe.coulomb(P.detach(), want_energy=True)
    J, EJ = _eri_j_raw(P.detach(), eri.detach(), want_energy=True)
    return J, EJ[0]


def xc_point_fused(name: str, clip: float, coef: Sequence[float], rho, grad_rho, tau, lapl, ehf, weights):
    """(E_xc, rho_bar, grad_rho_bar, tau_bar, lapl_bar, ehf_bar) of a closed-form functional with the constant coefficient row
    `coef` in one pass per grid point (gdft_xc_point_fused); no autograd -- the predictor's first-order path."""
    import ctypes

    L = lib()
    pid = _lib.PW_IDS[name]
    rho = _c(rho.detach())
    N = int(rho.shape[0])
    grad_rho = _c(grad_rho.detach()) if grad_rho is not None else None
    tau = _c(tau.detach()) if tau is not None else None
    lapl = _c(lapl.detach()) if lapl is not None else None
    ehf = _c(ehf.detach()) if ehf is not None else None
    W = int(ehf.shape[0]) if ehf is not None else 0
    weights = _c(weights.detach())
    dev = rho.device
    E = torch.empty((1,), dtype=F64, device=dev)
    rb = torch.empty_like(rho)
    gb = torch.empty_like(grad_rho) if grad_rho is not None else None
    tb = torch.empty_like(tau) if tau is not None else None
    lb = torch.empty_like(lapl) if lapl is not None else None
    eb = torch.empty_like(ehf) if ehf is not None else None
    ws = workspace(L.gdft_xc_point_workspace(N), dev)
    carr = (ctypes.c_double * len(coef))(*[float(c) for c in coef])
    check(L.gdft_xc_point_fused(stream_ptr(), N, pid, float(clip), carr, len(coef), ptr(rho), ptr(grad_rho), ptr(tau), ptr(lapl), ptr(ehf), W,
                                ptr(weights), ptr(E), ptr(rb), ptr(gb), ptr(tb), ptr(lb), ptr(eb), wptr(ws), ws.numel()), "gdft_xc_point_fused")
    return E[0], rb, gb, tb, lb, eb


def nonxc_energy(P: torch.Tensor, h1e: torch.Tensor, J: torch.Tensor, nuclear_repulsion) -> torch.Tensor:
    """E_nuc + <P, h1e> + 1/2 <P, J> (grad_dft/molecule.py:697-733) as one kernel; no autograd (the predictor's first-order path)."""
    P, h1e, J = _c(P.detach()), _c(h1e.detach()), _c(J.detach())
    n = int(P.shape[0])
    if not isinstance(nuclear_repulsion, torch.Tensor):
        nuclear_repulsion = torch.tensor(float(nuclear_repulsion), dtype=F64, device=P.device)
    en = _c(nuclear_repulsion.detach().reshape(1).to(P.device))
    out = torch.empty((1,), dtype=F64, device=P.device)
    check(lib().gdft_nonxc_energy(stream_ptr(), n, ptr(P), ptr(h1e), ptr(J), ptr(en), ptr(out)), "gdft_nonxc_energy")
    return out[0]


# ---------------------------------------------------------------------------------------------------------
# packed rep_tensor: the pair-symmetric quarter of the ERI sweep, laid out once per molecule (like the packed basis)
# ---------------------------------------------------------------------------------------------------------
ERI_SYMMETRY_RTOL = 1e-13  # max |asymmetry| / max |value| below which a rep_tensor counts as pair-symmetric


def eri_symmetry_defect(eri_rows: torch.Tensor, n: int, row0: int = 0) -> Tuple[float, float]:
    """(max |(pq|rt) - (pq|tr)| and |(pq|rt) - (qp|rt)| where both rows are present, max |(pq|rt)|) over the (p,q) rows
    [row0, row0 + rows) -- host-synchronous, a setup call."""
    L = lib()
    eri_rows = _c(eri_rows)
    rows = eri_rows.numel() // (n * n)
    out = torch.empty(2, dtype=F64, device=eri_rows.device)
    ws = workspace(L.gdft_eri_packed_workspace(n), eri_rows.device)
    check(L.gdft_eri_symmetry_defect(stream_ptr(), n, int(row0), rows, ptr(eri_rows), ptr(out), wptr(ws), ws.numel()), "gdft_eri_symmetry_defect")
    a, b = out.tolist()
    return a, b


class PackedERI:
    """Pair rows [pair0, pair0 + pairs) of a pair-symmetric rep_tensor as packed[pairs][npair_pad] (include/gdft_b200.h,
    "packed rep_tensor").  `complete` = all n(n+1)/2 pair rows are here (single-GPU molecule)."""

    def __init__(self, packed: torch.Tensor, n: int, pair0: int, pairs: int, exchange_symmetric: bool = False):
        self.packed, self.n, self.pair0, self.pairs = packed, int(n), int(pair0), int(pairs)
        self.npair = int(lib().gdft_eri_npair(n))
        self.complete = pair0 == 0 and pairs == self.npair
        self.exchange_symmetric = exchange_symmetric  # (pq|rt) == (rt|pq) verified: the transposed sweep is the sweep itself

    @staticmethod
    def _alloc(n: int, pairs: int, device) -> torch.Tensor:
        return torch.empty(int(lib().gdft_eri_packed_bytes(n, pairs)) // 8, dtype=F64, device=device)

    @classmethod
    def from_rows(cls, eri_rows: torch.Tensor, n: int, row0: int = 0) -> "PackedERI":
        """From a contiguous block of (p,q) rows (the whole tensor: row0 = 0, n*n rows): the pair rows whose (p >= q) source
        row lies inside the block -- a contiguous pair range, because pair(i,j) -> i*n + j is increasing."""
        eri_rows = _c(eri_rows)
        rows = eri_rows.numel() // (n * n)
        r1 = row0 + rows

        def first_pair_at_or_after(r):  # smallest pair index whose row i*n+j (j <= i) is >= r
            i, j = divmod(r, n)
            if i >= n:
                return n * (n + 1) // 2
            return i * (i + 1) // 2 + j if j <= i else (i + 1) * (i + 2) // 2

        pair0, pair1 = first_pair_at_or_after(row0), first_pair_at_or_after(r1)
        pairs = pair1 - pair0
        packed = cls._alloc(n, max(pairs, 1), eri_rows.device)
        if pairs > 0:
            check(lib().gdft_eri_pack(stream_ptr(), n, 0, int(row0), rows, ptr(eri_rows), pair0, pairs, ptr(packed)), "gdft_eri_pack")
        return cls(packed, n, pair0, pairs)

    @classmethod
    def from_pair_rows(cls, block: torch.Tensor, n: int, pair0: int) -> "PackedERI":
        """From a block that holds exactly the (p >= q) rows of the pairs [pair0, pair0 + block.shape[0]) (balanced sharding)."""
        block = _c(block)
        pairs = block.numel() // (n * n)
        packed = cls._alloc(n, max(pairs, 1), block.device)
        if pairs > 0:
            check(lib().gdft_eri_pack(stream_ptr(), n, 1, 0, pairs, ptr(block), int(pair0), pairs, ptr(packed)), "gdft_eri_pack")
        return cls(packed, n, pair0, pairs)

    def coulomb(self, P: torch.Tensor, want_energy: bool = False):
        """J[n,n] of the local pair rows (zero elsewhere unless complete) and, optionally, 1/2 <P, J>."""
        L = lib()
        P = _c(P)
        n = self.n
        if tuple(P.shape) != (n, n):
            raise TypeError(f"rdm1 must be [{n}, {n}], got {tuple(P.shape)}")
        J = torch.empty((n, n), dtype=F64, device=P.device)
        EJ = torch.empty((1,), dtype=F64, device=P.device) if want_energy else None
        ws = workspace(L.gdft_eri_packed_workspace(n), P.device)
        with _timed("gdft_eri_jk"):
            check(L.gdft_eri_j_packed(stream_ptr(), n, self.pair0, self.pairs, ptr(self.packed), ptr(P), ptr(J), ptr(EJ), wptr(ws), ws.numel()),
                  "gdft_eri_j_packed")
        return (J, EJ[0]) if want_energy else J


class _CoulombJPacked(Function):
    """J through the packed sweep.  With the exchange symmetry (pq|rt) = (rt|pq) verified the operator is self-adjoint and
    its VJP is the same call; otherwise the transposed sweep of the original tensor supplies it."""

    @staticmethod
    def forward(ctx, P, pe, eri):
        ctx.pe, ctx.eri = pe, eri
        return pe.coulomb(P)

    @staticmethod
    def backward(ctx, Jbar):
        pe = ctx.pe
        if not pe.complete:
            raise NotImplementedError("the VJP of a partial (sharded) packed sweep is not bound: the sharded predictor is first order")
        if pe.exchange_symmetric:
            return _CoulombJPacked.apply(Jbar, pe, ctx.eri), None, None
        return _CoulombJT.apply(Jbar, ctx.eri), None, None


_PACKED_ERI: list = []
PACK_ERI_MIN_N = 16


def _cache_entry(cache: list, tensor: torch.Tensor, extra: tuple, limit: int = 6) -> dict:
    """The cache entry of THIS tensor object at its current version counter (weak reference: an address reused by another
    tensor after this one died can never hit), created on first sight; the oldest entries are dropped beyond `limit`."""
    import weakref

    cache[:] = [e for e in cache if e["ref"]() is not None]
    for e in cache:
        if e["ref"]() is tensor and e["version"] == tensor._version and e["extra"] == extra:
            return e
    e = {"ref": weakref.ref(tensor), "version": tensor._version, "extra": extra, "uses": 0, "packed": None}
    cache.append(e)
    del cache[:-limit]
    return e


def packed_eri_for(eri: torch.Tensor, count_use: bool = True) -> Optional[PackedERI]:
    """The packed form of a FULL rep_tensor [n,n,n,n], or None.  Policy (env GDFT_PACK_ERI = auto | always | never): `auto`
    packs a tensor at its SECOND use (a one-off J pays one sweep, an SCF loop or a training run pays a quarter of a sweep
    per build from then on); packing checks the pair symmetries on the device first (two extra sweeps, host-synchronous,
    once per tensor) and is skipped for tensors that are not symmetric to ERI_SYMMETRY_RTOL.  Never packs during CUDA-graph
    capture.  Keyed by the tensor OBJECT (weak reference) and its version counter: in-place edits invalidate the packed copy,
    and a new tensor that happens to reuse a dead tensor's address can never pick it up."""
    import os

    mode = os.environ.get("GDFT_PACK_ERI", "auto")
    if mode == "never" or eri is None or not eri.is_cuda or eri.dim() != 4 or eri.dtype != F64:
        return None
    n = int(eri.shape[-1])
    if tuple(eri.shape) != (n, n, n, n) or n < PACK_ERI_MIN_N:
        return None
    entry = _cache_entry(_PACKED_ERI, eri, ())
    if entry["packed"] is not None:
        return entry["packed"] or None  # False: examined and refused
    if count_use:
        entry["uses"] += 1
    if (mode != "always" and entry["uses"] < 2) or torch.cuda.is_current_stream_capturing():
        return None
    try:
        e = _c(eri.detach())
        asym, big = eri_symmetry_defect(e, n, 0)
        if not (asym <= ERI_SYMMETRY_RTOL * big):
            entry["packed"] = False
            return None
        pe = PackedERI.from_rows(e, n, 0)
        # exchange symmetry of the packed matrix itself: (ij|kl) == (kl|ij)
        sq = pe.packed.view(pe.pairs, -1)[:, :pe.npair]
        pe.exchange_symmetric = bool(float((sq - sq.T).abs().max()) <= ERI_SYMMETRY_RTOL * big) if pe.npair <= 8192 else _exchange_symmetric_blocked(sq, big)
        entry["packed"] = pe
        return pe
    except torch.OutOfMemoryError:
        entry["packed"] = False
        return None


def release_packed_eri() -> None:
    """Drop every cached packed rep_tensor (they are otherwise dropped when their source tensor dies and the cache is next
    consulted)."""
    from . import distributed as gdist

    _PACKED_ERI.clear()
    gdist._PACKED_BLOCKS.clear()


def _exchange_symmetric_blocked(sq: torch.Tensor, big: float, block: int = 4096) -> bool:
    m = sq.shape[0]
    worst = 0.0
    for i in range(0, m, block):
        for j in range(i, m, block):
            worst = max(worst, float((sq[i:i + block, j:j + block] - sq[j:j + block, i:i + block].T).abs().max()))
    return worst <= ERI_SYMMETRY_RTOL * big


def coulomb_j_auto(P: torch.Tensor, eri: torch.Tensor) -> torch.Tensor:
    """`coulomb_j` through the packed sweep when the tensor has one (see `packed_eri_for`), else the plain sweep.  What the
    predictor and the `Molecule` methods call; `coulomb_j` itself always sweeps the tensor as given."""
    pe = packed_eri_for(eri)
    if pe is not None:
        return _CoulombJPacked.apply(P, pe, eri)
    return _CoulombJ.apply(P, eri)


# ---------------------------------------------------------------------------------------------------------
# XC quadrature
# ---------------------------------------------------------------------------------------------------------
class _XCIntegrate(Function):
    @staticmethod
    def forward(ctx, c, d, w, clip):
        L = lib()
        c, d, w = _c(c), _c(d), _c(w)
        N, F = int(d.shape[0]), int(d.shape[1])
        if c.dim() != 2 or c.shape[1] != F or c.shape[0] not in (1, N) or tuple(w.shape) != (N,):
            raise TypeError(f"shapes: coefficients {tuple(c.shape)}, densities {tuple(d.shape)}, weights {tuple(w.shape)}")
        E = torch.empty((1,), dtype=F64, device=d.device)
        ws = workspace(L.gdft_workspace_bytes(_lib.OP_XC_INTEGRATE, N, 0, 0, 0), d.device)
        check(L.gdft_xc_integrate_fwd(stream_ptr(), N, F, int(c.shape[0]), ptr(c), ptr(d), ptr(w), float(clip), ptr(E), wptr(ws), ws.numel()),
              "gdft_xc_integrate_fwd")
        ctx.save_for_backward(c, d, w)
        ctx.clip = float(clip)
        return E[0]

    @staticmethod
    def backward(ctx, Ebar):
        c, d, w = ctx.saved_tensors
        clip = ctx.clip
        if torch.is_grad_enabled():
            # higher-order request (create_graph=True): express the VJP with differentiable torch ops
            e = (c * d).sum(dim=1)
            wc = torch.where(w.abs() > clip, w, torch.zeros_like(w))
            eb = torch.where(e.abs() > clip, Ebar * wc, torch.zeros_like(wc)).unsqueeze(1)
            cbar = eb * d
            if c.shape[0] == 1:
                cbar = cbar.sum(dim=0, keepdim=True)
            return (cbar if ctx.needs_input_grad[0] else None), (eb * c if ctx.needs_input_grad[1] else None), None, None
        L = lib()
        N, F = int(d.shape[0]), int(d.shape[1])
        cbar = torch.empty_like(c) if ctx.needs_input_grad[0] else None
        dbar = torch.empty_like(d) if ctx.needs_input_grad[1] else None
        Eb = _c(Ebar.reshape(1))
        ws = workspace(L.gdft_workspace_bytes(_lib.OP_XC_INTEGRATE, N, 0, 0, 0), d.device)
        check(L.gdft_xc_integrate_bwd(stream_ptr(), N, F, int(c.shape[0]), ptr(c), ptr(d), ptr(w), clip, ptr(Eb), ptr(cbar), ptr(dbar),
                                      wptr(ws), ws.numel()), "gdft_xc_integrate_bwd")
        return cbar, dbar, None, None


def xc_integrate(coefficients: torch.Tensor, densities: torch.Tensor, weights: torch.Tensor, clip: float = 1e-30) -> torch.Tensor:
    """E = sum_r aclip(w_r) aclip(aclip(sum_f c[r,f] d[r,f]))."""
    return _XCIntegrate.apply(coefficients, densities, weights, clip)


# ---------------------------------------------------------------------------------------------------------
# closed-form per-point features
# ---------------------------------------------------------------------------------------------------------
class _Pointwise(Function):
    @staticmethod
    def forward(ctx, pw_id, clip, rho, grho, tau, lapl):
        L = lib()
        rho, grho, tau, lapl = _c(rho), _c(grho), _c(tau), _c(lapl)
        N = int(rho.shape[0])
        F = int(L.gdft_pointwise_ncols(pw_id))
        out = torch.empty((N, F), dtype=F64, device=rho.device)
        check(L.gdft_pointwise_fwd(stream_ptr(), N, pw_id, float(clip), ptr(rho), ptr(grho), ptr(tau), ptr(lapl), ptr(out)),
              "gdft_pointwise_fwd")
        ctx.pw_id, ctx.clip = pw_id, float(clip)
        ctx.present = (grho is not None, tau is not None, lapl is not None)
        ctx.save_for_backward(*[t for t in (rho, grho, tau, lapl) if t is not None])
        return out

    @staticmethod
    def backward(ctx, out_bar):
        saved = list(ctx.saved_tensors)
        rho = saved.pop(0)
        grho = saved.pop(0) if ctx.present[0] else None
        tau = saved.pop(0) if ctx.present[1] else None
        lapl = saved.pop(0) if ctx.present[2] else None
        rb, gb, tb, lb = _PointwiseVJP.apply(ctx.pw_id, ctx.clip, rho, grho, tau, lapl, out_bar)
        need = ctx.needs_input_grad
        return None, None, rb if need[2] else None, gb if need[3] else None, tb if need[4] else None, lb if need[5] else None


class _PointwiseVJP(Function):
    """(rho, grad_rho, tau, lapl, out_bar) -> input cotangents (gdft_pointwise_bwd); differentiable once more through
    gdft_pointwise_bwd2, which is what training through the SCF loop (grad of a function of V_xc) needs."""

    @staticmethod
    def forward(ctx, pw_id, clip, rho, grho, tau, lapl, out_bar):
        L = lib()
        out_bar = _c(out_bar)
        N = int(rho.shape[0])
        rb = torch.empty_like(rho)
        gb = torch.empty_like(grho) if grho is not None else None
        tb = torch.empty_like(tau) if tau is not None else None
        lb = torch.empty_like(lapl) if lapl is not None else None
        check(L.gdft_pointwise_bwd(stream_ptr(), N, pw_id, clip, ptr(rho), ptr(grho), ptr(tau), ptr(lapl), ptr(out_bar),
                                   ptr(rb), ptr(gb), ptr(tb), ptr(lb)), "gdft_pointwise_bwd")
        ctx.pw_id, ctx.clip = pw_id, clip
        ctx.present = (grho is not None, tau is not None, lapl is not None)
        ctx.save_for_backward(*[t for t in (rho, grho, tau, lapl, out_bar) if t is not None])
        ctx.set_materialize_grads(False)
        return rb, gb, tb, lb

    @staticmethod
    @once_differentiable  # third order is not bound: differentiating through this raises instead of silently dropping terms
    def backward(ctx, u_rho, u_grho, u_tau, u_lapl):
        L = lib()
        saved = list(ctx.saved_tensors)
        rho = saved.pop(0)
        grho = saved.pop(0) if ctx.present[0] else None
        tau = saved.pop(0) if ctx.present[1] else None
        lapl = saved.pop(0) if ctx.present[2] else None
        out_bar = saved.pop(0)
        N = int(rho.shape[0])
        need = ctx.needs_input_grad
        rt = torch.empty_like(rho) if need[2] else None
        gt = torch.empty_like(grho) if (grho is not None and need[3]) else None
        tt = torch.empty_like(tau) if (tau is not None and need[4]) else None
        lt = torch.empty_like(lapl) if (lapl is not None and need[5]) else None
        obb = torch.empty_like(out_bar) if need[6] else None
        check(L.gdft_pointwise_bwd2(stream_ptr(), N, ctx.pw_id, ctx.clip, ptr(rho), ptr(grho), ptr(tau), ptr(lapl), ptr(out_bar),
                                    ptr(_c(u_rho)), ptr(_c(u_grho)), ptr(_c(u_tau)), ptr(_c(u_lapl)),
                                    ptr(obb), ptr(rt), ptr(gt), ptr(tt), ptr(lt)), "gdft_pointwise_bwd2")
        return None, None, rt, gt, tt, lt, obb


def pointwise(name: str, rho, grad_rho=None, tau=None, lapl=None, clip: float = 1e-30) -> torch.Tensor:
    """out[N,F] of one closed-form feature set (see _lib.PW_IDS); differentiable to second order."""
    return _Pointwise.apply(_lib.PW_IDS[name], clip, rho, grad_rho, tau, lapl)


# ---------------------------------------------------------------------------------------------------------
# coefficient-network residual block (row f2)
# ---------------------------------------------------------------------------------------------------------
LN_ELU_MAX_WIDTH = 512


class _ResidualLayerNormElu(Function):
    """elu(LayerNorm(y + ybias + res) * scale + bias); ybias (the Dense bias) may be None."""

    @staticmethod
    def forward(ctx, y, ybias, res, scale, bias, eps):
        L = lib()
        y, ybias, res, scale, bias = _c(y), _c(ybias), _c(res), _c(scale), _c(bias)
        N, W = int(y.shape[0]), int(y.shape[1])
        out = torch.empty_like(y)
        stats = torch.empty((N, 2), dtype=F64, device=y.device)
        with _timed("gdft_ln_elu_fwd"):
            check(L.gdft_dense_ln_elu_fwd(stream_ptr(), N, W, ptr(y), ptr(ybias), ptr(res), ptr(scale), ptr(bias), float(eps), ptr(out),
                                          ptr(stats)), "gdft_dense_ln_elu_fwd")
        ctx.save_for_backward(y, ybias, res, scale, bias, stats, out)  # `out` is the next layer's saved input anyway
        ctx.eps = float(eps)
        return out

    @staticmethod
    @once_differentiable  # second order goes through the composite path, chosen up front by the caller
    def backward(ctx, out_bar):
        y, ybias, res, scale, bias, stats, fwd_out = ctx.saved_tensors
        L = lib()
        N, W = int(y.shape[0]), int(y.shape[1])
        need = ctx.needs_input_grad
        zbar = torch.empty_like(y)
        ybbar = torch.empty_like(ybias) if (ybias is not None and need[1]) else None
        sbar = torch.empty_like(scale) if need[3] else None
        bbar = torch.empty_like(bias) if need[4] else None
        pg = ybbar is not None or sbar is not None or bbar is not None
        ws = workspace(L.gdft_workspace_bytes(_lib.OP_LN_ELU, N, W, 0, 0), y.device) if pg else None
        with _timed("gdft_ln_elu_bwd"):
            check(L.gdft_dense_ln_elu_bwd(stream_ptr(), N, W, ptr(y), ptr(ybias), ptr(res), ptr(scale), ptr(bias), ptr(stats),
                                          ptr(fwd_out), ptr(_c(out_bar)), ptr(zbar), ptr(sbar), ptr(bbar), ptr(ybbar), wptr(ws),
                                          ws.numel() if ws is not None else 0), "gdft_dense_ln_elu_bwd")
        return (zbar if need[0] else None), ybbar, (zbar if (res is not None and need[2]) else None), sbar, bbar, None


class first_order_build:
    """Context in which the predictor promises that nothing evaluated inside will be differentiated twice
    (`xc_energy_and_grads` without create_graph): first-order-only fused kernels may be used."""

    depth = 0

    def __init__(self, active: bool = True):
        self.active = active

    def __enter__(self):
        if self.active:
            first_order_build.depth += 1

    def __exit__(self, *exc):
        if self.active:
            first_order_build.depth -= 1
        return False


def residual_layernorm_elu_supported(y: torch.Tensor) -> bool:
    return (first_order_build.depth > 0 and y.is_cuda and y.dtype == F64 and y.dim() == 2 and y.shape[1] % 2 == 0
            and y.shape[1] <= LN_ELU_MAX_WIDTH)


def residual_layernorm_elu(y: torch.Tensor, res: torch.Tensor, scale: torch.Tensor, bias: torch.Tensor, eps: float = 1e-6,
                           ybias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """elu(LayerNorm(y + ybias + res) * scale + bias): one fused pass forward, one reverse (first order).  `ybias` is the
    bias of the Dense layer that produced y (pass the bare GEMM output as y): its broadcast add and the column-sum of
    its cotangent ride in the same two passes.  The caller picks the host-framework composite instead when a second
    derivative will be taken through it."""
    return _ResidualLayerNormElu.apply(y, ybias, res, scale, bias, eps)


# ---------------------------------------------------------------------------------------------------------
# coefficient-network Dense layers as FP64 tensor-core GEMMs of the library (row f2; csrc/dense_gemm.cu)
# ---------------------------------------------------------------------------------------------------------
DENSE_MAX_WIDTH = 256
# Reverse pass of a trunk, per block boundary: False = the plain GEMM (x_bar = z_bar K^T + z_bar, at cuBLAS parity: 2.00 ms
# at N = 5e5, W = 256) followed by the streaming ELU/LayerNorm reverse of the previous block (0.86 ms, 4.75 TB/s); True = ONE
# GEMM whose epilogue does that reverse (gdft_dense_block_bwd).  The fused form saves 3 GB of HBM traffic per block but
# measures 3.19 ms against 2.86: its epilogue is a chain of dependent global loads on a CTA that holds 128 registers of
# accumulators, and the co-resident CTA alone cannot keep the tensor pipe full meanwhile (ncu: tensor pipe 91 % active in
# the plain kernel, 64 % in the fused one).  Kept selectable (and tested); the forward fusion wins (2.51 vs 2.61 ms) and is used.
DENSE_BWD_CHAIN = False


def _dense_ws(N: int, K: int, Wd: int, device) -> torch.Tensor:
    return workspace(lib().gdft_workspace_bytes(_lib.OP_DENSE, N, K, Wd, 0), device)


def _pad_cols(t: torch.Tensor, width: int) -> torch.Tensor:
    if t.shape[1] == width:
        return t
    out = t.new_zeros((t.shape[0], width))
    out[:, :t.shape[1]] = t
    return out


def _up8(k: int) -> int:
    return (k + 7) // 8 * 8


def _dense_fwd_raw(x: torch.Tensor, kernel_t: torch.Tensor, bias: Optional[torch.Tensor], res: Optional[torch.Tensor]) -> torch.Tensor:
    """x[N,K] @ kernel_t[Wd,K]^T (+ bias) (+ res): shapes already legal for the kernel (K even, Wd % 8 == 0, Wd <= 256)."""
    N, K = int(x.shape[0]), int(x.shape[1])
    Wd = int(kernel_t.shape[0])
    out = torch.empty((N, Wd), dtype=F64, device=x.device)
    with _timed("gdft_dense_fwd"):
        check(lib().gdft_dense_fwd(stream_ptr(), N, K, Wd, ptr(x), ptr(kernel_t), ptr(bias), ptr(res), ptr(out)), "gdft_dense_fwd")
    return out


def _dense_bwd_weight_raw(x: torch.Tensor, z_bar: torch.Tensor) -> torch.Tensor:
    N, K, Wd = int(x.shape[0]), int(x.shape[1]), int(z_bar.shape[1])
    out = torch.empty((K, Wd), dtype=F64, device=x.device)
    ws = _dense_ws(N, K, Wd, x.device)
    with _timed("gdft_dense_bwd_weight"):
        check(lib().gdft_dense_bwd_weight(stream_ptr(), N, K, Wd, ptr(x), ptr(z_bar), ptr(out), wptr(ws), ws.numel()), "gdft_dense_bwd_weight")
    return out


def dense_supported(x: torch.Tensor, kernel: torch.Tensor) -> bool:
    """Whether `x @ kernel + bias` goes through the library GEMM: a first-order build on float64 CUDA tensors, output width
    up to 256 (after padding both widths to multiples of 8: 11 -> 16 inputs, 3 -> 8 outputs for DM21's first and last layer)."""
    return (first_order_build.depth > 0 and x.is_cuda and x.dtype == F64 and x.dim() == 2 and kernel.dim() == 2 and x.shape[0] > 0
            and x.shape[1] == kernel.shape[0] and _up8(kernel.shape[1]) <= DENSE_MAX_WIDTH and _up8(kernel.shape[0]) <= DENSE_MAX_WIDTH)


class _DenseLayer(Function):
    """flax Dense, y = x K + k (grad_dft/functional.py:803,811,414), on gdft_dense_fwd; input, kernel and bias cotangents from
    gdft_dense_fwd (x_bar = y_bar K^T) and gdft_dense_bwd_weight (K_bar = x^T y_bar).  Widths are zero-padded to multiples of 8."""

    @staticmethod
    def forward(ctx, x, kernel, bias):
        x, kernel = _c(x), _c(kernel)
        K, Wd = int(kernel.shape[0]), int(kernel.shape[1])
        K8, W8 = _up8(K), _up8(Wd)
        xp = _pad_cols(x, K8)
        kp = kernel.new_zeros((K8, W8))
        kp[:K, :Wd] = kernel
        bp = None
        if bias is not None:
            bp = bias.new_zeros(W8)
            bp[:Wd] = bias
        out = _dense_fwd_raw(xp, kp.t().contiguous(), bp, None)
        ctx.save_for_backward(xp, kp)
        ctx.dims = (K, Wd, bias is not None)
        return out if W8 == Wd else out[:, :Wd]

    @staticmethod
    @once_differentiable
    def backward(ctx, y_bar):
        xp, kp = ctx.saved_tensors
        K, Wd, has_bias = ctx.dims
        yb = _pad_cols(_c(y_bar), kp.shape[1])
        need = ctx.needs_input_grad
        x_bar = k_bar = b_bar = None
        if need[0]:
            x_bar = _dense_fwd_raw(yb, kp, None, None)  # y_bar K^T: the transposed right operand is K as stored
            x_bar = x_bar if x_bar.shape[1] == K else x_bar[:, :K]
        if need[1]:
            k_bar = _dense_bwd_weight_raw(xp, yb)[:K, :Wd]
        if has_bias and need[2]:
            b_bar = y_bar.sum(dim=0)
        return x_bar, k_bar, b_bar


def dense_layer(x: torch.Tensor, kernel: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    return _DenseLayer.apply(x, kernel, bias)


def residual_trunk_supported(x: torch.Tensor, width: int) -> bool:
    return (first_order_build.depth > 0 and x.is_cuda and x.dtype == F64 and x.dim() == 2 and x.shape[0] > 0 and x.shape[1] == width
            and width % 8 == 0 and width <= DENSE_MAX_WIDTH)


class _ResidualTrunk(Function):
    """A run of residual blocks x <- elu(LayerNorm(x K + k + x) * scale + bias) (the loop of DM21's default_nn,
    grad_dft/functional.py:809-819) as ONE differentiable unit: forward = one fused GEMM kernel per block
    (gdft_dense_block_fwd); reverse = one streaming pass for the last block's ELU/LayerNorm (gdft_dense_block_bwd_last), then
    per block one GEMM whose epilogue undoes the previous block's ELU/LayerNorm (gdft_dense_block_bwd) and one split-K GEMM
    for the kernel cotangent (gdft_dense_bwd_weight).  params = (kernel, dense_bias, scale, bias) per block."""

    @staticmethod
    def forward(ctx, x, eps, *params):
        L_ = lib()
        x = _c(x)
        N, W = int(x.shape[0]), int(x.shape[1])
        nb = len(params) // 4
        params = tuple(_c(t) for t in params)
        outs, xhats, rstds = [x], [], []
        for l in range(nb):
            kernel, kb, scale, bias = params[4 * l:4 * l + 4]
            out = torch.empty_like(x)
            xhat = torch.empty_like(x)
            rstd = torch.empty((N,), dtype=F64, device=x.device)
            kt = kernel.t().contiguous()
            with _timed("gdft_dense_block_fwd"):
                check(L_.gdft_dense_block_fwd(stream_ptr(), N, W, ptr(outs[-1]), ptr(kt), ptr(kb), ptr(scale), ptr(bias), float(eps), ptr(out),
                                              ptr(xhat), ptr(rstd)), "gdft_dense_block_fwd")
            outs.append(out)
            xhats.append(xhat)
            rstds.append(rstd)
        ctx.nb = nb
        ctx.save_for_backward(*outs, *xhats, *rstds, *params)
        return outs[-1]

    @staticmethod
    @once_differentiable  # second order goes through the composite path, chosen up front by the caller
    def backward(ctx, out_bar):
        L_ = lib()
        nb = ctx.nb
        saved = ctx.saved_tensors
        outs, xhats, rstds = saved[:nb + 1], saved[nb + 1:2 * nb + 1], saved[2 * nb + 1:3 * nb + 1]
        params = saved[3 * nb + 1:]
        N, W = int(outs[0].shape[0]), int(outs[0].shape[1])
        dev = outs[0].device
        ws = _dense_ws(N, W, W, dev)
        grads = [None] * (4 * nb)

        def pgrads(l):
            need = ctx.needs_input_grad[2 + 4 * l:2 + 4 * l + 4]
            kb = torch.empty((W,), dtype=F64, device=dev) if need[1] else None
            sb = torch.empty((W,), dtype=F64, device=dev) if need[2] else None
            bb = torch.empty((W,), dtype=F64, device=dev) if need[3] else None
            grads[4 * l + 1], grads[4 * l + 2], grads[4 * l + 3] = kb, sb, bb
            return kb, sb, bb

        z_bar = torch.empty_like(outs[0])
        kb, sb, bb = pgrads(nb - 1)
        with _timed("gdft_dense_block_bwd_last"):
            check(L_.gdft_dense_block_bwd_last(stream_ptr(), N, W, ptr(_c(out_bar)), ptr(outs[nb]), ptr(xhats[nb - 1]), ptr(rstds[nb - 1]),
                                               ptr(params[4 * (nb - 1) + 2]), ptr(z_bar), ptr(sb), ptr(bb), ptr(kb), wptr(ws), ws.numel()),
                  "gdft_dense_block_bwd_last")
        for l in range(nb - 1, -1, -1):
            kernel = params[4 * l]
            if ctx.needs_input_grad[2 + 4 * l]:
                grads[4 * l] = _dense_bwd_weight_raw(outs[l], z_bar)  # K_bar = x_l^T z_bar_l
            if l == 0:
                break
            prev = torch.empty_like(z_bar)
            kb, sb, bb = pgrads(l - 1)
            if DENSE_BWD_CHAIN:
                with _timed("gdft_dense_block_bwd"):
                    check(L_.gdft_dense_block_bwd(stream_ptr(), N, W, ptr(z_bar), ptr(kernel), ptr(outs[l]), ptr(xhats[l - 1]), ptr(rstds[l - 1]),
                                                  ptr(params[4 * (l - 1) + 2]), ptr(prev), ptr(sb), ptr(bb), ptr(kb), wptr(ws), ws.numel()),
                          "gdft_dense_block_bwd")
            else:
                x_bar_l = _dense_fwd_raw(z_bar, kernel, None, z_bar)  # cotangent of block l-1's output: z_bar K^T + z_bar
                with _timed("gdft_dense_block_bwd_last"):
                    check(L_.gdft_dense_block_bwd_last(stream_ptr(), N, W, ptr(x_bar_l), ptr(outs[l]), ptr(xhats[l - 1]), ptr(rstds[l - 1]),
                                                       ptr(params[4 * (l - 1) + 2]), ptr(prev), ptr(sb), ptr(bb), ptr(kb), wptr(ws), ws.numel()),
                          "gdft_dense_block_bwd_last")
            z_bar = prev
        x_bar = None
        if ctx.needs_input_grad[0]:
            x_bar = _dense_fwd_raw(z_bar, params[0], None, z_bar)  # z_bar K_0^T + z_bar (the residual branch)
        return (x_bar, None, *grads)


def residual_trunk(x: torch.Tensor, blocks, eps: float = 1e-6) -> torch.Tensor:
    """`blocks`: sequence of (kernel[W,W], dense_bias[W], scale[W], bias[W]); see _ResidualTrunk."""
    flat = [t for blk in blocks for t in blk]
    return _ResidualTrunk.apply(x, eps, *flat)


# ---------------------------------------------------------------------------------------------------------
# SCF harness: small symmetric eigenproblem (row f1)
# ---------------------------------------------------------------------------------------------------------
def sym_eigh_supported(A: torch.Tensor) -> bool:
    return A.is_cuda and A.dtype == F64 and A.dim() >= 2 and A.shape[-1] == A.shape[-2] and 0 < A.shape[-1] <= lib().gdft_sym_eigh_max_n()


def sym_eigh(A: torch.Tensor, V0: Optional[torch.Tensor] = None, info: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """(eigenvalues ascending [..., n], eigenvectors as columns [..., n, n]) of symmetric A[..., n, n]; no autograd
    (evaluate.safe_eigh supplies the VJP), no host synchronisation.  `V0` (same shape, ORTHOGONAL: the eigenvectors of a
    nearby matrix) warm-starts the Jacobi sweeps.  `info` (optional int32 CUDA tensor, one entry per matrix) receives the
    sweep count, -1 if the sweep bound was hit, -2 for a non-finite result -- written stream-ordered, read it when convenient."""
    A = _c(A.detach())
    n = int(A.shape[-1])
    batch = A.numel() // (n * n)
    evals = torch.empty(A.shape[:-1], dtype=F64, device=A.device)
    evecs = torch.empty_like(A)
    if V0 is not None:
        V0 = _c(V0.detach())
        if V0.shape != A.shape:
            raise TypeError(f"V0 {tuple(V0.shape)} does not match A {tuple(A.shape)}")
    info_p = None
    if info is not None:
        if info.dtype != torch.int32 or not info.is_cuda or info.numel() < batch or not info.is_contiguous():
            raise TypeError("info must be a contiguous int32 CUDA tensor with one entry per matrix")
        from ctypes import c_void_p
        info_p = c_void_p(info.data_ptr())
    with _timed("gdft_sym_eigh"):
        check(lib().gdft_sym_eigh_ex(stream_ptr(), batch, n, ptr(A), ptr(V0), ptr(evals), ptr(evecs), info_p), "gdft_sym_eigh_ex")
    return evals, evecs


# ---------------------------------------------------------------------------------------------------------
# SCF harness: the n x n tail of a DIIS iteration of a small molecule as two kernels (row f1)
# ---------------------------------------------------------------------------------------------------------
def scf_stage_supported(t: torch.Tensor, max_diis: int) -> bool:
    import os

    if os.environ.get("GDFT_SCF_FUSED", "1") == "0":  # A/B switch: the host-framework tail
        return False
    return t.is_cuda and t.dtype == F64 and t.shape[-1] <= lib().gdft_scf_stage_max_n() and max_diis <= 16


def aufbau_occupations(evals: torch.Tensor, occ_prev: torch.Tensor) -> torch.Tensor:
    """Aufbau occupations (grad_dft/molecule.py:851-889) by stable rank counting: one small kernel instead of a sort, a
    scatter and four elementwise launches; no autograd (occupations are piecewise constant)."""
    evals, occ_prev = _c(evals.detach()), _c(occ_prev.detach())
    n = int(evals.shape[-1])
    occ = torch.empty_like(evals)
    check(lib().gdft_aufbau_occupations(stream_ptr(), n, ptr(evals), ptr(occ_prev), ptr(occ)), "gdft_aufbau_occupations")
    return occ


def scf_diis_step(cycle: int, fock, rdm1, overlap, L_inv, fock_vec, err_vec, gram):
    """One CDIIS step of grad_dft/evaluate.py:1111-1205 on the loop-private ring buffers (updated in place) followed by the
    Cholesky reduction of eigenproblem.py:125-127: returns (C = L^-1 F' L^-T, F', x) -- see gdft_scf_diis_step."""
    n, m = int(fock.shape[-1]), int(fock_vec.shape[0])
    fock, rdm1, overlap, L_inv = _c(fock.detach()), _c(rdm1.detach()), _c(overlap.detach()), _c(L_inv.detach())
    C = torch.empty_like(fock)
    fock_out = torch.empty_like(fock)
    x = torch.empty((2, m), dtype=F64, device=fock.device)
    check(lib().gdft_scf_diis_step(stream_ptr(), n, m, int(cycle), ptr(fock), ptr(rdm1), ptr(overlap), ptr(L_inv), ptr(fock_vec), ptr(err_vec),
                                   ptr(gram), ptr(x), ptr(fock_out), ptr(C)), "gdft_scf_diis_step")
    return C, fock_out, x


def scf_occupy(evals, V, L_inv, occ_prev):
    """(mo_coeff = L^-T V, aufbau occupations, rdm1 = C occ C^T) -- eigenproblem.py:129, molecule.py:815-889."""
    n = int(V.shape[-1])
    evals, V, L_inv, occ_prev = _c(evals.detach()), _c(V.detach()), _c(L_inv.detach()), _c(occ_prev.detach())
    mo_coeff, rdm1 = torch.empty_like(V), torch.empty_like(V)
    mo_occ = torch.empty_like(evals)
    check(lib().gdft_scf_occupy(stream_ptr(), n, ptr(evals), ptr(V), ptr(L_inv), ptr(occ_prev), ptr(mo_coeff), ptr(mo_occ), ptr(rdm1)), "gdft_scf_occupy")
    return mo_coeff, mo_occ, rdm1


class _AbsClip(Function):
    """out = where(|src| > thr, x, 0): abs_clip for x = src, and the same mask applied to a cotangent in its VJP (which is
    again this Function, so any order of differentiation closes over the one kernel).  No gradient flows to `src`."""

    @staticmethod
    def forward(ctx, x, src, thr):
        x, src = _c(x), _c(src)
        out = torch.empty_like(x)
        check(lib().gdft_abs_clip(stream_ptr(), x.numel(), ptr(x), ptr(src), float(thr), ptr(out)), "gdft_abs_clip")
        ctx.save_for_backward(src)
        ctx.thr = float(thr)
        return out

    @staticmethod
    def backward(ctx, g):
        (src,) = ctx.saved_tensors
        return _AbsClip.apply(g, src, ctx.thr), None, None


def abs_clip(arr: torch.Tensor, threshold: float) -> torch.Tensor:
    """grad_dft/molecule.py:687-689 as one kernel (value and VJP)."""
    return _AbsClip.apply(arr, arr.detach(), threshold)


def diis_gram(err_vec: torch.Tensor) -> torch.Tensor:
    """einsum("iskl,jskl->sij", err_vec, err_vec) for the CDIIS ring buffer err_vec[m, 2, n, n] (no autograd)."""
    e = _c(err_vec.detach())
    m, n = int(e.shape[0]), int(e.shape[-1])
    gram = torch.empty((2, m, m), dtype=F64, device=e.device)
    check(lib().gdft_diis_gram(stream_ptr(), m, n, ptr(e), ptr(gram)), "gdft_diis_gram")
    return gram


def diis_matrix(err_vec: torch.Tensor, cycle: int) -> torch.Tensor:
    """The bordered CDIIS matrix B[2, m+1, m+1] (grad_dft/evaluate.py:1167-1181) from the ring buffer err_vec[m, 2, n, n]."""
    e = _c(err_vec.detach())
    m, n = int(e.shape[0]), int(e.shape[-1])
    B = torch.empty((2, m + 1, m + 1), dtype=F64, device=e.device)
    check(lib().gdft_diis_matrix(stream_ptr(), m, n, int(cycle), ptr(e), ptr(B)), "gdft_diis_matrix")
    return B


def diis_combine(x: torch.Tensor, fock_vec: torch.Tensor) -> torch.Tensor:
    """einsum("si,isjk->sjk", x, fock_vec) for x[2, m] and the ring buffer fock_vec[m, 2, n, n] (no autograd)."""
    x, f = _c(x.detach()), _c(fock_vec.detach())
    m, n = int(f.shape[0]), int(f.shape[-1])
    out = torch.empty((2, n, n), dtype=F64, device=f.device)
    check(lib().gdft_diis_combine(stream_ptr(), m, n, ptr(x), ptr(f), ptr(out)), "gdft_diis_combine")
    return out


# ---------------------------------------------------------------------------------------------------------
# chi generation tail (row f4)
# ---------------------------------------------------------------------------------------------------------
def chi_contract_(chi: torch.Tensor, start: int, w: int, ao: torch.Tensor, rdm1: torch.Tensor, nu: torch.Tensor) -> torch.Tensor:
    """chi[start:start+Nc, w] = einsum("sbd,rb,rda->rsa", rdm1, ao[start:start+Nc], nu) written in place into the
    reference-layout tensor chi[N, W, 2, n] (grad_dft/interface/pyscf.py:1110-1124); nu[Nc, n, n] is one _nu_chunk."""
    if chi.dtype != F64 or not chi.is_contiguous() or chi.dim() != 4:
        raise TypeError("chi must be a contiguous float64 [N, W, 2, n] tensor")
    ao, rdm1, nu = _c(ao), _c(rdm1.detach()), _c(nu)
    N, W, two, n = (int(x) for x in chi.shape)
    Nc = int(nu.shape[0])
    if two != 2 or tuple(nu.shape) != (Nc, n, n) or tuple(rdm1.shape) != (2, n, n) or tuple(ao.shape) != (N, n):
        raise TypeError(f"shape mismatch: chi {tuple(chi.shape)}, nu {tuple(nu.shape)}, rdm1 {tuple(rdm1.shape)}, ao {tuple(ao.shape)}")
    if not (0 <= start and start + Nc <= N and 0 <= w < W):
        raise IndexError("chunk outside the grid / omega outside the list")
    if Nc == 0:
        return chi
    ptr(chi), ptr(ao)  # device / dtype / contiguity checks
    from ctypes import c_void_p
    ao_p = c_void_p(ao.data_ptr() + 8 * start * n)
    chi_p = c_void_p(chi.data_ptr() + 8 * ((start * W + w) * 2 * n))
    with _timed("gdft_chi_contract"):
        check(lib().gdft_chi_contract(stream_ptr(), Nc, n, ao_p, n, ptr(rdm1), ptr(nu), chi_p, W * 2 * n), "gdft_chi_contract")
    return chi


# ---------------------------------------------------------------------------------------------------------
# predictor glue (no autograd: these sit after value_and_grad in grad_dft/train.py:148-215)
# ---------------------------------------------------------------------------------------------------------
def fock_assemble(h1e, J, rdm1_bar, clip: float = 1e-30) -> torch.Tensor:
    L = lib()
    h1e, J, rdm1_bar = _c(h1e), _c(J), _c(rdm1_bar)
    n = int(h1e.shape[0])
    fock = torch.empty((2, n, n), dtype=F64, device=h1e.device)
    check(L.gdft_fock_assemble(stream_ptr(), n, ptr(h1e), ptr(J), ptr(rdm1_bar), float(clip), ptr(fock)), "gdft_fock_assemble")
    return fock


def fock_add_sym_(fock, V, clip: float = 1e-30) -> torch.Tensor:
    L = lib()
    V = _c(V)
    n = int(fock.shape[1])
    check(L.gdft_fock_add_sym(stream_ptr(), n, ptr(V), float(clip), ptr(fock)), "gdft_fock_add_sym")
    return fock
