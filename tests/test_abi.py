"""The C-ABI shared library loads without a GPU and exports every symbol include/gdft_b200.h declares; argument
validation that is decided on the host (before any launch) returns the documented status codes."""
import ctypes
import re
from pathlib import Path

import pytest

from graddft_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "gdft_b200.h").read_text()


def declared_symbols():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(gdft_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    if not _lib.LIB_PATH.exists():
        from graddft_b200.build import build
        build()
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/gdft_b200.h but not exported"
    # the ctypes table mirrors the header one to one
    assert sorted(_lib.SIGNATURES) == syms


def test_host_side_validation_without_gpu():
    L = _lib.lib()
    assert L.gdft_version() >= 100
    assert L.gdft_npad(43) == 48 and L.gdft_npad(400) == 400
    assert L.gdft_packed_basis_bytes(1000, 43, 4) == 4 * 1000 * 48 * 8
    assert L.gdft_pointwise_ncols(_lib.PW_IDS["B3LYP_SET"]) == 4 and L.gdft_pointwise_ncols(_lib.PW_IDS["DM21_INPUTS"]) == 7
    assert L.gdft_workspace_bytes(_lib.OP_DENSITY_FWD, 1000, 43, 0, 0) >= 2 * 48 * 48 * 8
    assert L.gdft_workspace_bytes(_lib.OP_DENSITY_BWD, 1000, 43, 0, 0) > 0
    assert L.gdft_status_string(1) == b"bad shape" and L.gdft_status_string(3) == b"workspace too small"
    # shape / argument errors are reported before anything touches the device
    assert L.gdft_density_fwd(None, -1, 4, 1, 1, None, None, None, 0, None, None, None, None, None, None, 0) == 1
    assert L.gdft_density_fwd(None, 10, 4, 0, 1, None, None, None, 0, None, None, None, None, None, None, 0) == 5
    assert L.gdft_density_fwd(None, 10, 4, 1, 1, None, None, None, 0, None, None, None, None, None, None, 0) == 5  # NULL packed
    assert L.gdft_eri_jk(None, 0, None, None, None, None, None, None, 0) == 1
    assert L.gdft_pointwise_fwd(None, 10, 99, 1e-30, None, None, None, None, None) == 5
    assert L.gdft_xc_integrate_fwd(None, 10, 40, 1, None, None, None, 1e-30, None, None, 0) == 1


def test_ops_refuse_cpu_tensors():
    import torch
    from graddft_b200 import ops

    with pytest.raises(_lib.GdftError):
        ops.PackedBasis(torch.zeros(4, 3, dtype=torch.float64))


def test_xla_adapter_dims_layout():
    from graddft_b200 import jax_ffi

    jax_ffi.check_layout()
    assert len(jax_ffi.pack_dims(N=5, n=3)) == _lib.lib().gdft_xla_dims_size()


def test_fused_xc_rows_are_the_functionals_own_coefficient_rows():
    """popular_functionals.fused_xc_spec (host logic of the one-pass per-point kernel): for every closed-form functional the
    constant row handed to gdft_xc_point_fused is what `Functional.coefficients_for` broadcasts over the columns of the
    functional's own feature set (grad_dft/functional.py:246-250: einsum "rf,rf->r" with a [1, 1] or [1, F] row), the
    exact-exchange column last; functionals outside the table take the generic chain."""
    import torch

    import graddft_b200 as gd
    from graddft_b200 import _lib
    from graddft_b200.popular_functionals import fused_xc_spec

    L = _lib.lib()
    for fun in (gd.LSDA, gd.B88, gd.VWN, gd.LYP, gd.PW92, gd.B3LYP):
        name, row, omegas = fused_xc_spec(fun)
        F = int(L.gdft_pointwise_ncols(_lib.PW_IDS[name]))
        ncols = F + (1 if omegas else 0)
        assert len(row) == ncols
        like = torch.empty((5, ncols), dtype=torch.float64)
        want = fun.coefficients_for(None, None, like)
        assert tuple(want.shape) == (1, ncols)
        assert torch.equal(want.cpu().reshape(-1), torch.tensor(row, dtype=torch.float64))
    assert fused_xc_spec(gd.DM21()) is None
