"""ctypes binding of the C-ABI in include/gdft_b200.h.  No arithmetic here.

The library is loaded lazily from the package directory (built in-tree by graddft_b200.build).  There
is no CPU implementation behind these symbols and no fallback: a missing library, a missing symbol
or a non-zero status raises.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p
from pathlib import Path

import torch

LIB_PATH = Path(__file__).resolve().parent / "libgdft_b200.so"

GDFT_RHO, GDFT_GRAD, GDFT_TAU, GDFT_LAPL, GDFT_HF = 1, 2, 4, 8, 16
OP_DENSITY_FWD, OP_DENSITY_BWD, OP_HF_FOCK, OP_ERI_J, OP_XC_INTEGRATE, OP_LN_ELU, OP_DENSE = 1, 2, 3, 4, 5, 6, 7
PW_IDS = {
    "LSDA_X": 0, "B88_X": 1, "VWN_C": 2, "LYP_C": 3, "PW92_C": 4, "B3LYP_SET": 5, "B88_SET": 6,
    "DM21_INPUTS": 7, "DM21_LDA": 8, "DM21_GGA": 9, "DM21_MGGA": 10, "FEAT_LDA": 11, "FEAT_GGA": 12, "FEAT_MGGA": 13,
}

# name -> (restype, argtypes); mirrors include/gdft_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "gdft_version": (c_int, []),
    "gdft_last_cuda_error": (c_int, []),
    "gdft_launch_count": (ctypes.c_ulonglong, []),
    "gdft_status_string": (c_char_p, [c_int]),
    "gdft_device_supported": (c_int, []),
    "gdft_npad": (c_int64, [c_int64]),
    "gdft_packed_basis_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "gdft_pack_basis": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, c_int]),
    "gdft_pack_chi": (c_int, [_P, c_int64, c_int64, c_int, _P, _P]),
    "gdft_workspace_bytes": (c_size_t, [c_int, c_int64, c_int64, c_int, c_int]),
    "gdft_density_fwd": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_density_bwd": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_hf_fock": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_hf_fock_sum": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_eri_jk": (c_int, [_P, c_int64, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_eri_k_transpose": (c_int, [_P, c_int64, _P, _P, _P, _P, c_size_t]),
    "gdft_eri_j_transpose": (c_int, [_P, c_int64, _P, _P, _P, _P, c_size_t]),
    "gdft_eri_j_rows": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "gdft_eri_j_transpose_rows": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, c_size_t]),
    "gdft_xc_integrate_fwd": (c_int, [_P, c_int64, c_int, c_int64, _P, _P, _P, c_double, _P, _P, c_size_t]),
    "gdft_xc_integrate_bwd": (c_int, [_P, c_int64, c_int, c_int64, _P, _P, _P, c_double, _P, _P, _P, _P, c_size_t]),
    "gdft_pointwise_ncols": (c_int, [c_int]),
    "gdft_pointwise_fwd": (c_int, [_P, c_int64, c_int, c_double, _P, _P, _P, _P, _P]),
    "gdft_pointwise_bwd": (c_int, [_P, c_int64, c_int, c_double, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gdft_pointwise_bwd2": (c_int, [_P, c_int64, c_int, c_double] + [_P] * 14),
    "gdft_ln_elu_fwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, c_double, _P, _P]),
    "gdft_ln_elu_bwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_dense_ln_elu_fwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, c_double, _P, _P]),
    "gdft_dense_ln_elu_bwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_dense_supported": (c_int, [c_int64, c_int64]),
    "gdft_dense_fwd": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "gdft_dense_block_fwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, c_double, _P, _P, _P]),
    "gdft_dense_block_bwd": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_dense_bwd_weight": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, c_size_t]),
    "gdft_dense_block_bwd_last": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_sym_eigh_max_n": (c_int, []),
    "gdft_sym_eigh": (c_int, [_P, c_int64, c_int64, _P, _P, _P]),
    "gdft_sym_eigh_warm": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P]),
    "gdft_sym_eigh_ex": (c_int, [_P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "gdft_abs_clip": (c_int, [_P, c_int64, _P, _P, c_double, _P]),
    "gdft_diis_gram": (c_int, [_P, c_int, c_int64, _P, _P]),
    "gdft_diis_matrix": (c_int, [_P, c_int, c_int64, c_int, _P, _P]),
    "gdft_diis_combine": (c_int, [_P, c_int, c_int64, _P, _P, _P]),
    "gdft_chi_contract_max_n": (c_int64, []),
    "gdft_chi_contract": (c_int, [_P, c_int64, c_int64, _P, c_int64, _P, _P, _P, c_int64]),
    "gdft_fock_assemble": (c_int, [_P, c_int64, _P, _P, _P, c_double, _P]),
    "gdft_fock_add_sym": (c_int, [_P, c_int64, _P, c_double, _P]),
    "gdft_aufbau_occupations": (c_int, [_P, c_int64, _P, _P, _P]),
    "gdft_scf_stage_max_n": (c_int, []),
    "gdft_scf_diis_step": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gdft_scf_occupy": (c_int, [_P, c_int64, _P, _P, _P, _P, _P, _P, _P]),
    "gdft_xc_point_workspace": (c_size_t, [c_int64]),
    "gdft_xc_point_fused": (c_int, [_P, c_int64, c_int, c_double, ctypes.POINTER(c_double), c_int, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_nonxc_energy": (c_int, [_P, c_int64, _P, _P, _P, _P, _P]),
    "gdft_eri_npair": (c_int64, [c_int64]),
    "gdft_eri_packed_bytes": (c_size_t, [c_int64, c_int64]),
    "gdft_eri_packed_workspace": (c_size_t, [c_int64]),
    "gdft_eri_symmetry_defect": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, c_size_t]),
    "gdft_eri_pack": (c_int, [_P, c_int64, c_int, c_int64, c_int64, _P, c_int64, c_int64, _P]),
    "gdft_eri_j_packed": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, c_size_t]),
    "gdft_nccl_available": (c_int, []),
    "gdft_nccl_unique_id_bytes": (c_size_t, []),
    "gdft_nccl_unique_id": (c_int, [_P]),
    "gdft_nccl_comm_create": (c_int, [_P, c_int, c_int, ctypes.POINTER(_P)]),
    "gdft_nccl_comm_destroy": (c_int, [_P]),
    "gdft_allreduce_fock": (c_int, [_P, _P, _P, c_size_t]),
    "gdft_comm_handle_bytes": (c_size_t, []),
    "gdft_comm_create": (c_int, [c_int, c_int, c_size_t, ctypes.POINTER(_P)]),
    "gdft_comm_buffer": (_P, [_P]),
    "gdft_comm_capacity": (c_size_t, [_P]),
    "gdft_comm_handle": (c_int, [_P, _P]),
    "gdft_comm_connect": (c_int, [_P, _P]),
    "gdft_comm_connect_local": (c_int, [_P, ctypes.POINTER(_P)]),
    "gdft_comm_status": (c_int, [_P, ctypes.POINTER(c_int), ctypes.POINTER(ctypes.c_ulonglong)]),
    "gdft_comm_destroy": (c_int, [_P]),
    "gdft_allreduce_fock_p2p": (c_int, [_P, _P, c_size_t]),
    "gdft_xla_last_status": (c_int, []),
    "gdft_xla_dims_size": (c_size_t, []),
}
# XLA custom-call adapters: void(stream, void** buffers, const char* opaque, size_t opaque_len)
for _name in ("pack_basis", "pack_chi", "density_fwd", "density_bwd", "hf_fock", "eri_j", "eri_j_transpose", "xc_integrate_fwd", "xc_integrate_bwd",
              "pointwise_fwd", "pointwise_bwd", "pointwise_bwd2", "eri_j_rows", "eri_j_transpose_rows", "ln_elu_fwd", "ln_elu_bwd",
            "dense_ln_elu_fwd", "dense_ln_elu_bwd", "sym_eigh", "chi_contract", "diis_gram", "diis_combine", "eri_jk", "eri_k_transpose"):
    SIGNATURES[f"gdft_{_name}_xla"] = (None, [_P, ctypes.POINTER(_P), c_char_p, c_size_t])

_lib = None


class GdftError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """The loaded C-ABI library; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise GdftError(
                f"{LIB_PATH} is missing: build it with `python -m graddft_b200.build` "
                "(graddft_b200 has no CPU or eager fallback for its kernels)"
            )
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        L = lib()
        msg = L.gdft_status_string(status).decode()
        extra = f" (cuda error {L.gdft_last_cuda_error()})" if status == 4 else ""
        raise GdftError(f"{what}: {msg}{extra}")


def ptr(t: torch.Tensor | None):
    """Device pointer of a contiguous float64 CUDA tensor (or NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GdftError("graddft_b200 kernels need CUDA tensors (there is no CPU path)")
    if t.dtype != torch.float64:
        raise TypeError(f"expected float64, got {t.dtype}")
    if not t.is_contiguous():
        raise GdftError("tensor must be contiguous")
    _same_device(t)
    return c_void_p(t.data_ptr())


def wptr(t: torch.Tensor | None):
    """Device pointer of a raw (uint8) workspace tensor (or NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise GdftError("graddft_b200 kernels need CUDA tensors (there is no CPU path)")
    _same_device(t)
    return c_void_p(t.data_ptr())


def _same_device(t: torch.Tensor) -> None:
    """The C entry points launch on the CURRENT device's current stream (stream_ptr) and never switch devices, so a
    tensor living on another GPU would be dereferenced by a kernel running on the wrong device: refuse it here."""
    cur = torch.cuda.current_device()
    if t.device.index != cur:
        raise GdftError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: wrap the call in "
                        f"`with torch.cuda.device({t.device.index}):` (kernels launch on the current device's current stream)")


def stream_ptr() -> c_void_p:
    """cudaStream_t of the current device's current stream; `ptr`/`wptr` check that every tensor lives on that device."""
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
