"""Grid sharding across the GPUs of one box (not in the reference, which is single-device).

Every grid op on the path is a map over grid rows followed by a sum over rows, so the rows of
ao / grad_ao / grad_n_ao / chi / weights are split into contiguous blocks, one per rank (one process per GPU);
rdm1, params and the n x n matrices are replicated.  Each rank produces a partial E_xc and a partial V_xc
(and partial explicit-HF Fock terms); ONE all-reduce of the packed buffer [E_xc | V_xc(2,n,n)] per XC build
finishes the job (payload 8*(2n^2+1) bytes: 2.56 MB at n=400, latency-bound on NVLink 5).  No other
collective exists on the path.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(N: int, rank: int, world: int, align: int = 128) -> Tuple[int, int]:
    """Contiguous row block [lo, hi) of rank `rank`: blocks are multiples of `align` rows (the CTA row tile of the
    density kernel) except the last, and differ by at most one tile."""
    tiles = (N + align - 1) // align
    base, extra = divmod(tiles, world)
    lo_t = rank * base + min(rank, extra)
    hi_t = lo_t + base + (1 if rank < extra else 0)
    return min(N, lo_t * align), min(N, hi_t * align)


_ROW_FIELDS = ("ao", "grad_ao", "grad_n_ao2", "chi", "weights", "coords")


def pair_bounds(n: int, rank: int, world: int, align: int = 32) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the n(n+1)/2 (p >= q) pair rows of a pair-symmetric rep_tensor for rank `rank`."""
    return shard_bounds(n * (n + 1) // 2, rank, world, align=align)


def pair_row_indices(n: int, lo: int, hi: int, device=None) -> torch.Tensor:
    """Flattened (p,q) row index p*n + q (p >= q) of the pairs lo .. hi-1 in pair order pair(p,q) = p(p+1)/2 + q."""
    pr = torch.arange(lo, hi, dtype=torch.int64, device=device)
    p = ((torch.sqrt(8.0 * pr.to(torch.float64) + 1.0) - 1.0) * 0.5).floor().to(torch.int64)
    p = torch.where(p * (p + 1) // 2 > pr, p - 1, p)
    p = torch.where((p + 1) * (p + 2) // 2 <= pr, p + 1, p)
    return p * n + (pr - p * (p + 1) // 2)


def shard_molecule_tensors(mol: Dict[str, torch.Tensor], rank: int, world: int, shard_eri=False) -> Dict[str, torch.Tensor]:
    """The rank's row block of every grid-sized tensor; everything else is passed through (replicated).  With
    `shard_eri` the rows of rep_tensor are split as well (at n = 400 the tensor is 205 GB and cannot be replicated):
    True / "rows": `rep_tensor` becomes the contiguous block [rows, n, n] of (p,q) rows and `eri_row0` its first row;
    "pairs": `rep_tensor` becomes the block of (p >= q) rows of this rank's share of the n(n+1)/2 pair rows and `eri_pair0`
    its first pair -- half the rows, evenly balanced, for pair-symmetric tensors (what the packed sweep needs)."""
    N = int(mol["weights"].shape[0])
    lo, hi = shard_bounds(N, rank, world)
    out = dict(mol)
    for k in _ROW_FIELDS:
        if out.get(k) is not None:
            out[k] = out[k][lo:hi].contiguous()
    if shard_eri and out.get("rep_tensor") is not None:
        eri = out["rep_tensor"]
        n = int(eri.shape[-1])
        if shard_eri == "pairs":
            p0, p1 = pair_bounds(n, rank, world)
            out["rep_tensor"] = eri.reshape(n * n, n, n)[pair_row_indices(n, p0, p1, eri.device)].contiguous()
            out["eri_pair0"] = p0
        else:
            r0, r1 = shard_bounds(n * n, rank, world, align=32)  # 32 rows = one CTA pass of the sweep kernel
            out["rep_tensor"] = eri.reshape(n * n, n, n)[r0:r1].contiguous()
            out["eri_row0"] = r0
    return out


@dataclass(frozen=True)
class GridShard:
    """How a `Molecule` is spread over the process group: this rank holds grid rows [lo, hi) of every grid-sized
    tensor and, when `eri_row0` is not None, the (p,q) rows [eri_row0, eri_row0 + rep_tensor.shape[0]) of
    rep_tensor.  `energy_predictor` reads it from the molecule and closes each Fock build with one all-reduce."""

    group: Any
    rank: int
    world: int
    eri_row0: Optional[int] = None
    eri_pair0: Optional[int] = None  # rep_tensor holds the (p >= q) rows of the pairs [eri_pair0, eri_pair0 + rep_tensor.shape[0])

    @property
    def eri_sharded(self) -> bool:
        return self.eri_row0 is not None or self.eri_pair0 is not None


def attach_shard(molecule, shard: "GridShard"):
    """Mark `molecule` (already holding this rank's rows) as one shard of a grid-sharded molecule.  The mark
    survives `Molecule.replace`."""
    object.__setattr__(molecule, "_shard", shard)
    return molecule


def shard_molecule(mol: Dict[str, torch.Tensor], rank: int, world: int, device=None, group=None, shard_eri: bool = False):
    """`Molecule` holding rank `rank`'s shard of the tensor dict `mol` (keys as `molecule_from_tensors` takes them),
    marked so that `energy_predictor` / the SCF loops all-reduce each Fock build over `group`."""
    from .molecule import molecule_from_tensors

    part = shard_molecule_tensors(mol, rank, world, shard_eri=shard_eri)
    return attach_shard(molecule_from_tensors(part, device), GridShard(group, rank, world, part.get("eri_row0"), part.get("eri_pair0")))


_PACKED_BLOCKS: list = []


def _packed_block(rep_tensor: torch.Tensor, n: int, shard: "GridShard"):
    """Packed form of this rank's rep_tensor block (pair-symmetric columns; for a (p,q)-row block also only its p >= q rows),
    built at the block's second use like `ops.packed_eri_for`; None when packing is off, refused or not yet due."""
    import os
    from . import ops

    mode = os.environ.get("GDFT_PACK_ERI", "auto")
    if mode == "never" or not rep_tensor.is_cuda or n < ops.PACK_ERI_MIN_N:
        return None
    entry = ops._cache_entry(_PACKED_BLOCKS, rep_tensor, (shard.eri_row0, shard.eri_pair0), limit=4)
    if entry["packed"] is not None:
        return entry["packed"] or None
    entry["uses"] += 1
    if (mode != "always" and entry["uses"] < 2) or torch.cuda.is_current_stream_capturing():
        return None
    try:
        if shard.eri_pair0 is not None:
            # pair rows: only the column symmetry can be checked locally (row symmetry is the caller's statement)
            rows = int(rep_tensor.shape[0])
            col = rep_tensor - rep_tensor.transpose(1, 2) if rows * n * n <= (1 << 28) else None
            ok = True if col is None else bool(float(col.abs().max()) <= ops.ERI_SYMMETRY_RTOL * float(rep_tensor.abs().max()))
            del col
            entry["packed"] = ops.PackedERI.from_pair_rows(rep_tensor, n, shard.eri_pair0) if ok else False
        else:
            asym, big = ops.eri_symmetry_defect(rep_tensor, n, shard.eri_row0)
            entry["packed"] = ops.PackedERI.from_rows(rep_tensor, n, shard.eri_row0) if asym <= ops.ERI_SYMMETRY_RTOL * big else False
    except torch.OutOfMemoryError:
        entry["packed"] = False
    return entry["packed"] or None


def local_coulomb(P: torch.Tensor, rep_tensor: torch.Tensor, shard: "GridShard") -> torch.Tensor:
    """J as this rank can compute it: the full matrix when rep_tensor is replicated, else the entries of its own rows
    written into a zero matrix (the all-reduce that follows assembles the rest).  Row blocks of a pair-symmetric tensor
    go through the packed sweep from their second use on (a quarter of the bytes; see include/gdft_b200.h)."""
    from . import ops

    if not shard.eri_sharded:
        return ops.coulomb_j_auto(P, rep_tensor)
    n = int(P.shape[0])
    pe = _packed_block(rep_tensor, n, shard)
    if pe is not None:
        return pe.coulomb(P)
    if shard.eri_pair0 is not None:
        # un-packed pair rows: the plain row sweep, scattered to both triangles
        rows = int(rep_tensor.shape[0])
        idx = pair_row_indices(n, shard.eri_pair0, shard.eri_pair0 + rows, P.device)
        vals = ops.coulomb_j_rows(P, rep_tensor)
        J = torch.zeros(n * n, dtype=P.dtype, device=P.device)
        J[idx] = vals
        J[(idx % n) * n + idx // n] = vals
        return J.reshape(n, n)
    J = torch.zeros(n * n, dtype=P.dtype, device=P.device)
    rows = int(rep_tensor.shape[0])
    J[shard.eri_row0:shard.eri_row0 + rows] = ops.coulomb_j_rows(P, rep_tensor)
    return J.reshape(n, n)


class _DevicePointer:
    """Exposes raw device memory of the library to the host framework (no copy): __cuda_array_interface__ v2."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class FockComm:
    """The exchange step of the sharded path inside the library (`gdft_allreduce_fock*`, include/gdft_b200.h).

    backend "p2p": the library's own one-launch reduce-scatter + all-gather kernel over NVLink peer memory.  The payload
    (`self.buffer`, a float64 view of IPC-shared device memory owned by the communicator) is ordinary device memory:
    the density VJP's second-stage reduce and the other partial results are written straight into it
    (`ops.density_bwd_into`), so nothing is packed or concatenated before the exchange; the call is stream-ordered, needs no
    host thread and can be captured in a CUDA graph.  Results are bitwise identical on every rank.
    backend "nccl": in-place `ncclAllReduce` on a communicator the library creates from a broadcast unique id
    (`gdft_allreduce_fock`, the signature of SURVEY.md 8b), on the same payload view.
    One process per GPU; handles / ids travel through `torch.distributed` object collectives of `group` (setup only)."""

    def __init__(self, capacity: int, device, group=None, backend: str = "p2p"):
        from . import _lib

        L = _lib.lib()
        self.group, self.backend, self.device = group, backend, torch.device(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._L, self._comm, self._nccl = L, None, None
        import ctypes

        with torch.cuda.device(self.device):
            if backend == "p2p":
                comm = ctypes.c_void_p()
                _lib.check(L.gdft_comm_create(self.rank, self.world, int(capacity), ctypes.byref(comm)), "gdft_comm_create")
                self._comm = comm
                hb = int(L.gdft_comm_handle_bytes())
                mine = ctypes.create_string_buffer(hb)
                _lib.check(L.gdft_comm_handle(comm, mine), "gdft_comm_handle")
                handles = [None] * self.world
                dist.all_gather_object(handles, bytes(mine.raw), group=group)
                _lib.check(L.gdft_comm_connect(comm, ctypes.create_string_buffer(b"".join(handles), hb * self.world)), "gdft_comm_connect")
                self.capacity = int(L.gdft_comm_capacity(comm))
                self.buffer = torch.as_tensor(_DevicePointer(int(L.gdft_comm_buffer(comm)), self.capacity), device=self.device)
                dist.barrier(group=group)  # every rank has opened every peer before the first exchange
            elif backend == "nccl":
                if not L.gdft_nccl_available():
                    raise _lib.GdftError("libnccl.so.2 could not be loaded")
                nb = int(L.gdft_nccl_unique_id_bytes())
                ident = ctypes.create_string_buffer(nb)
                if self.rank == 0:
                    _lib.check(L.gdft_nccl_unique_id(ident), "gdft_nccl_unique_id")
                box = [bytes(ident.raw)]
                dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
                comm = ctypes.c_void_p()
                _lib.check(L.gdft_nccl_comm_create(ctypes.create_string_buffer(box[0], nb), self.rank, self.world, ctypes.byref(comm)), "gdft_nccl_comm_create")
                self._nccl = comm
                self.capacity = int(capacity + (capacity & 1))
                self.buffer = torch.zeros(self.capacity, dtype=torch.float64, device=self.device)
            else:
                raise ValueError(f"unknown backend {backend!r}")

    def allreduce(self, count: Optional[int] = None) -> torch.Tensor:
        """buffer[:count] <- sum over ranks, in place, on the current stream; returns the view."""
        from . import _lib

        count = self.capacity if count is None else int(count)
        if self._comm is not None:
            _lib.check(self._L.gdft_allreduce_fock_p2p(_lib.stream_ptr(), self._comm, count), "gdft_allreduce_fock_p2p")
        else:
            _lib.check(self._L.gdft_allreduce_fock(self._nccl, _lib.stream_ptr(), _lib.ptr(self.buffer), count), "gdft_allreduce_fock")
        return self.buffer[:count]

    def status(self) -> int:
        """Host-synchronous: 0 ok, 1 some exchange timed out waiting for a peer (p2p backend)."""
        if self._comm is None:
            return 0
        import ctypes

        st, ep = ctypes.c_int(0), ctypes.c_ulonglong(0)
        self._L.gdft_comm_status(self._comm, ctypes.byref(st), ctypes.byref(ep))
        return int(st.value)

    def close(self) -> None:
        if self._comm is not None:
            torch.cuda.synchronize(self.device)
            self._L.gdft_comm_destroy(self._comm)
            self._comm = None
        if self._nccl is not None:
            torch.cuda.synchronize(self.device)
            self._L.gdft_nccl_comm_destroy(self._nccl)
            self._nccl = None


_FOCK_COMMS: Dict[Any, "FockComm"] = {}


def fock_comm(capacity: int, device, group=None, backend: Optional[str] = None) -> Optional["FockComm"]:
    """The process-wide communicator of (group, device), created on first use and grown when a larger payload is asked for.
    Returns None when the exchange cannot run inside the library (CPU tensors / gloo tests): callers then fall back to
    `torch.distributed.all_reduce`, which is the host framework's collective, not a different numerical path."""
    import os

    backend = backend or os.environ.get("GDFT_ALLREDUCE", "p2p")
    if backend == "torch" or torch.device(device).type != "cuda" or not (dist.is_available() and dist.is_initialized()):
        return None
    if dist.get_backend(group) != "nccl":
        return None
    key = (id(group) if group is not None else None, torch.device(device).index, backend)
    c = _FOCK_COMMS.get(key)
    if c is None or c.capacity < capacity:
        if c is not None:
            c.close()
        c = _FOCK_COMMS[key] = FockComm(max(int(capacity), 1 << 20), device, group, backend)
    return c


def packed_layout(sizes: Sequence[int]) -> Tuple[list, int]:
    """Offsets of consecutive payload segments, each starting on an even element (16-byte aligned for the 128-bit peer
    loads/stores of the exchange kernel and the vectorised reduce epilogues); returns (offsets, total)."""
    offs, off = [], 0
    for k in sizes:
        offs.append(off)
        off += int(k) + (int(k) & 1)
    return offs, off


def exchange_is_capturable(device, group=None) -> bool:
    """True when `allreduce_sum_packed` on `device` runs as the library's peer-memory kernel (backend "p2p"): stream-ordered,
    no host thread, device-resident epochs -- the only exchange a CUDA graph of the sharded SCF loop may contain."""
    import os

    if os.environ.get("GDFT_ALLREDUCE", "p2p") != "p2p" or torch.device(device).type != "cuda":
        return False
    return bool(dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl")


def allreduce_sum_packed(tensors: Sequence[torch.Tensor], group=None, skip: Sequence[int] = ()):
    """Sum each tensor over the ranks of `group` with ONE exchange on a flat float64 payload; entries whose index is in
    `skip` are already complete on every rank and are passed through untouched.  On CUDA with an NCCL process group the
    exchange runs inside the library (`FockComm`): tensors that already live in their payload segment (see
    `ops.density_bwd_into`) are not copied, and the returned tensors are VIEWS of the payload, valid until the next exchange.
    Otherwise (CPU / gloo) the host framework's all_reduce is used."""
    idx = [i for i in range(len(tensors)) if i not in skip]
    if not idx or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(tensors)
    offs, total = packed_layout([tensors[i].numel() for i in idx])
    comm = fock_comm(total, tensors[idx[0]].device, group)
    out = list(tensors)
    if comm is not None:
        for i, off in zip(idx, offs):
            seg = comm.buffer[off:off + tensors[i].numel()]
            if tensors[i].data_ptr() != seg.data_ptr():
                seg.copy_(tensors[i].reshape(-1))
        comm.allreduce(total)
        for i, off in zip(idx, offs):
            out[i] = comm.buffer[off:off + tensors[i].numel()].reshape(tensors[i].shape)
        return out
    flat = torch.cat([tensors[i].reshape(-1) for i in idx])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for i in idx:
        k = tensors[i].numel()
        out[i] = flat[off:off + k].reshape(tensors[i].shape)
        off += k
    return out


def payload_segment(index: int, sizes: Sequence[int], device, group=None) -> Optional[torch.Tensor]:
    """The payload segment entry `index` of an `allreduce_sum_packed([...])` call with these element counts will occupy
    (None when the exchange does not run inside the library): hand it to `ops.density_bwd_into`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    offs, total = packed_layout(sizes)
    comm = fock_comm(total, device, group)
    return None if comm is None else comm.buffer[offs[index]:offs[index] + int(sizes[index])]


def pack_xc(exc: torch.Tensor, vxc: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[E_xc | V_xc.ravel()] as one contiguous float64 buffer."""
    n2 = vxc.numel()
    if out is None:
        out = torch.empty(1 + n2, dtype=vxc.dtype, device=vxc.device)
    out[0] = exc
    out[1:] = vxc.reshape(-1)
    return out


def unpack_xc(buf: torch.Tensor, shape) -> Tuple[torch.Tensor, torch.Tensor]:
    return buf[0], buf[1:].reshape(shape)


def allreduce_xc(exc: torch.Tensor, vxc: torch.Tensor, group=None, buf: Optional[torch.Tensor] = None):
    """Sum the rank-local partial (E_xc, V_xc) over the grid shards: one exchange on the payload [V_xc | E_xc]
    (V_xc first: it is the segment the density VJP writes in place, `payload_segment(0, [vxc.numel(), 1], ...)`)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return exc, vxc
    vxc, exc = allreduce_sum_packed([vxc, exc.reshape(1)], group=group)
    return exc.reshape(()), vxc


# ---------------------------------------------------------------------------------------------------------
# molecule (batch) sharding: independent molecules of a training batch, one gradient all-reduce per step
# (the reference loops serially, grad_dft/train.py:493-494,519-528)
# ---------------------------------------------------------------------------------------------------------
def shard_molecules(costs, rank: int, world: int):
    """Indices of the molecules rank `rank` evaluates: longest-processing-time-first greedy balance on the given
    per-molecule costs (N_i * n_i^2), deterministic, identical on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += costs[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def allreduce_gradients(grads, loss: Optional[torch.Tensor] = None, group=None):
    """Sum parameter gradients (and the loss) over the molecule shards with ONE collective on a flat buffer."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return grads, loss
    flat = torch.cat([g.reshape(-1) for g in grads] + ([loss.reshape(1)] if loss is not None else []))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, off = [], 0
    for g in grads:
        out.append(flat[off:off + g.numel()].reshape(g.shape))
        off += g.numel()
    return out, (flat[off] if loss is not None else None)
