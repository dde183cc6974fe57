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


def shard_molecule_tensors(mol: Dict[str, torch.Tensor], rank: int, world: int, shard_eri: bool = False) -> Dict[str, torch.Tensor]:
    """The rank's row block of every grid-sized tensor; everything else is passed through (replicated).  With
    `shard_eri` the (p,q) rows of rep_tensor are split as well: `rep_tensor` becomes the block [rows, n, n] and
    `eri_row0` its first row (at n = 400 the tensor is 205 GB and cannot be replicated)."""
    N = int(mol["weights"].shape[0])
    lo, hi = shard_bounds(N, rank, world)
    out = dict(mol)
    for k in _ROW_FIELDS:
        if out.get(k) is not None:
            out[k] = out[k][lo:hi].contiguous()
    if shard_eri and out.get("rep_tensor") is not None:
        eri = out["rep_tensor"]
        n = int(eri.shape[-1])
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
    return attach_shard(molecule_from_tensors(part, device), GridShard(group, rank, world, part.get("eri_row0")))


def local_coulomb(P: torch.Tensor, rep_tensor: torch.Tensor, shard: "GridShard") -> torch.Tensor:
    """J as this rank can compute it: the full matrix when rep_tensor is replicated, else its own (p,q) rows written
    into a zero matrix (the all-reduce that follows assembles the rest)."""
    from . import ops

    if shard.eri_row0 is None:
        return ops.coulomb_j(P, rep_tensor)
    n = int(P.shape[0])
    J = torch.zeros(n * n, dtype=P.dtype, device=P.device)
    rows = int(rep_tensor.shape[0])
    J[shard.eri_row0:shard.eri_row0 + rows] = ops.coulomb_j_rows(P, rep_tensor)
    return J.reshape(n, n)


def allreduce_sum_packed(tensors: Sequence[torch.Tensor], group=None, skip: Sequence[int] = ()):
    """Sum each tensor over the ranks of `group` with ONE collective on a flat float64 buffer; entries whose index is
    in `skip` are already complete on every rank and are passed through untouched."""
    idx = [i for i in range(len(tensors)) if i not in skip]
    if not idx or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(tensors)
    flat = torch.cat([tensors[i].reshape(-1) for i in idx])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, off = list(tensors), 0
    for i in idx:
        k = tensors[i].numel()
        out[i] = flat[off:off + k].reshape(tensors[i].shape)
        off += k
    return out


def pack_xc(exc: torch.Tensor, vxc: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[E_xc | V_xc.ravel()] as one contiguous float64 buffer (the all-reduce payload)."""
    n2 = vxc.numel()
    if out is None:
        out = torch.empty(1 + n2, dtype=vxc.dtype, device=vxc.device)
    out[0] = exc
    out[1:] = vxc.reshape(-1)
    return out


def unpack_xc(buf: torch.Tensor, shape) -> Tuple[torch.Tensor, torch.Tensor]:
    return buf[0], buf[1:].reshape(shape)


def allreduce_xc(exc: torch.Tensor, vxc: torch.Tensor, group=None, buf: Optional[torch.Tensor] = None):
    """Sum the rank-local partial (E_xc, V_xc) over the grid shards: one collective on the packed buffer."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return exc, vxc
    buf = pack_xc(exc, vxc, buf)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return unpack_xc(buf, vxc.shape)


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
