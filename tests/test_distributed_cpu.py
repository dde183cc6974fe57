"""Host logic of the grid-sharded path on CPU: shard bounds, payload packing, and the one all-reduce per build with
world_size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from graddft_b200 import distributed as gdist


def test_shard_bounds_cover_and_balance():
    for N in (1, 127, 128, 129, 34_000, 500_000, 2_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [gdist.shard_bounds(N, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == N
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 128 + 127


def test_shard_tensors_and_pack_roundtrip():
    N, n = 1000, 5
    mol = {"ao": torch.randn(N, n, dtype=torch.float64), "grad_ao": torch.randn(N, n, 3, dtype=torch.float64),
           "weights": torch.rand(N, dtype=torch.float64), "coords": torch.randn(N, 3, dtype=torch.float64),
           "rdm1": torch.randn(2, n, n, dtype=torch.float64)}
    parts = [gdist.shard_molecule_tensors(mol, r, 3) for r in range(3)]
    assert torch.equal(torch.cat([p["ao"] for p in parts]), mol["ao"])
    assert torch.equal(torch.cat([p["weights"] for p in parts]), mol["weights"])
    assert all(p["rdm1"] is mol["rdm1"] for p in parts)
    e, v = torch.tensor(-1.5, dtype=torch.float64), torch.randn(2, n, n, dtype=torch.float64)
    e2, v2 = gdist.unpack_xc(gdist.pack_xc(e, v), v.shape)
    assert float(e2) == float(e) and torch.equal(v2, v)


def _run_world(target, world, *extra, attempts=3):
    """Spawn `world` gloo ranks of `target(rank, world, port, q, *extra)` and collect one result per rank (sorted by rank).  The
    free-port probe can lose a race with another process and the first `import torch` of a spawned child can take a minute on
    a cold container, so a failed rendezvous is retried on a fresh port instead of failing the suite."""
    last = None
    for _ in range(attempts):
        s = socket.socket()
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
        s.close()
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        procs = [ctx.Process(target=target, args=(r, world, port, q, *extra)) for r in range(world)]
        [p.start() for p in procs]
        try:
            res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
            [p.join(timeout=120) for p in procs]
            if all(p.exitcode == 0 for p in procs):
                return res
            last = RuntimeError(f"rank exit codes {[p.exitcode for p in procs]}")
        except Exception as exc:  # queue.Empty: a rank never reported
            last = exc
        for p in procs:
            if p.is_alive():
                p.kill()
            p.join(timeout=30)
    raise AssertionError(f"world-size-{world} run of {target.__name__} failed {attempts} times: {last!r}")


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6
    g = torch.Generator().manual_seed(100 + rank)
    e = torch.randn((), generator=g, dtype=torch.float64)
    v = torch.randn(2, n, n, generator=g, dtype=torch.float64)
    e_sum, v_sum = gdist.allreduce_xc(e, v)
    q.put((rank, float(e), v.clone(), float(e_sum), v_sum.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_xc_world2_gloo():
    res = _run_world(_worker, 2)
    e_tot = res[0][1] + res[1][1]
    v_tot = res[0][2] + res[1][2]
    for r in res:
        assert abs(r[3] - e_tot) < 1e-15 and torch.allclose(r[4], v_tot, rtol=0, atol=1e-15)


def test_shard_molecules_balanced_and_complete():
    g = torch.Generator().manual_seed(1993)
    costs = [float((1e4 * (1 + 3 * torch.rand((), generator=g))) * (12 + int(88 * torch.rand((), generator=g))) ** 2) for _ in range(64)]
    for world in (1, 2, 4, 8):
        parts = [gdist.shard_molecules(costs, r, world) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(64))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= 1.15 * (sum(costs) / world)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7 + rank)
    grads = [torch.randn(3, 4, generator=g, dtype=torch.float64), torch.randn(5, generator=g, dtype=torch.float64)]
    loss = torch.randn((), generator=g, dtype=torch.float64)
    out, l = gdist.allreduce_gradients(grads, loss)
    q.put((rank, [x.clone() for x in grads], float(loss), [x.clone() for x in out], float(l)))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_gradients_world2_gloo():
    res = _run_world(_grad_worker, 2)
    for k in range(2):
        tot = res[0][1][k] + res[1][1][k]
        assert torch.allclose(res[0][3][k], tot, rtol=0, atol=1e-15) and torch.allclose(res[1][3][k], tot, rtol=0, atol=1e-15)
    assert abs(res[0][4] - (res[0][2] + res[1][2])) < 1e-15


def test_shard_eri_rows_cover():
    n = 7
    mol = {"weights": torch.rand(300, dtype=torch.float64), "ao": torch.randn(300, n, dtype=torch.float64),
           "rep_tensor": torch.randn(n, n, n, n, dtype=torch.float64)}
    parts = [gdist.shard_molecule_tensors(mol, r, 3, shard_eri=True) for r in range(3)]
    assert torch.equal(torch.cat([p["rep_tensor"] for p in parts]).reshape(n, n, n, n), mol["rep_tensor"])
    assert parts[0]["eri_row0"] == 0
    for a, b in zip(parts, parts[1:]):
        assert b["eri_row0"] == a["eri_row0"] + a["rep_tensor"].shape[0]


def _packed_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(11 + rank)
    ts = [torch.randn((), generator=g, dtype=torch.float64), torch.randn(2, 3, 3, generator=g, dtype=torch.float64),
          torch.full((3, 3), 5.0, dtype=torch.float64), torch.randn(2, 3, 3, generator=g, dtype=torch.float64)]
    out = gdist.allreduce_sum_packed(ts, skip=(2,))
    q.put((rank, [t.clone() for t in ts], [t.clone() for t in out]))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_sum_packed_world2_gloo():
    res = _run_world(_packed_worker, 2)
    for k in (0, 1, 3):
        tot = res[0][1][k] + res[1][1][k]
        for r in res:
            assert torch.allclose(r[2][k], tot, rtol=0, atol=1e-15)
    for r in res:  # the skipped entry is passed through untouched
        assert torch.equal(r[2][2], torch.full((3, 3), 5.0, dtype=torch.float64))


def _spin_split_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graddft_b200 import evaluate
    g = torch.Generator().manual_seed(5)  # the same replicated inputs on every rank
    n = 9
    F = torch.randn(2, n, n, generator=g, dtype=torch.float64)
    F = F + F.transpose(1, 2)
    S = torch.randn(n, n, generator=g, dtype=torch.float64)
    S = torch.eye(n, dtype=torch.float64) + 0.05 * (S + S.T)
    shard = gdist.GridShard(None, rank, world)
    with torch.no_grad():
        w, C = evaluate.safe_fock_solver(F, S, None, shard)     # spin blocks split between the ranks
        w0, C0 = evaluate.safe_fock_solver(F, S)                # every rank solves both
    q.put((rank, w.clone(), C.clone(), w0.clone(), C0.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_spin_split_fock_solver_gloo(world):
    """The sharded SCF iteration's eigensolve: rank 0 solves spin 0, rank 1 spin 1, one all-reduce with a single non-zero
    contributor per entry -- every rank ends with the bits the unsplit solver produces (host logic; gloo on CPU)."""
    res = _run_world(_spin_split_worker, world)
    for rank, w, C, w0, C0 in res:
        assert torch.equal(w, res[0][1]) and torch.equal(C, res[0][2])      # replicated bit for bit
        assert torch.allclose(w, w0, rtol=0, atol=1e-13)
        assert torch.allclose(C.abs(), C0.abs(), rtol=0, atol=1e-10)        # eigenvector signs are the solver's choice


def test_pair_row_sharding_partitions_the_lower_triangle():
    """distributed.pair_bounds / pair_row_indices: the (p >= q) rows of a pair-symmetric rep_tensor, split evenly."""
    from graddft_b200 import distributed as gdist

    for n, world in ((7, 2), (43, 3), (264, 8)):
        npair = n * (n + 1) // 2
        seen = []
        for r in range(world):
            lo, hi = gdist.pair_bounds(n, r, world)
            idx = gdist.pair_row_indices(n, lo, hi)
            seen.append(idx)
            assert hi - lo <= -(-npair // world) + 32
        allrows = torch.cat(seen)
        p, q = allrows // n, allrows % n
        assert allrows.numel() == npair and bool((p >= q).all())
        assert torch.equal(allrows, torch.sort(allrows).values) and allrows.unique().numel() == npair
    mol = {"weights": torch.ones(256, dtype=torch.float64), "rep_tensor": torch.arange(5 ** 4, dtype=torch.float64).reshape(5, 5, 5, 5)}
    part = gdist.shard_molecule_tensors(mol, 1, 2, shard_eri="pairs")
    lo, hi = gdist.pair_bounds(5, 1, 2)
    assert part["eri_pair0"] == lo and part["rep_tensor"].shape == (hi - lo, 5, 5)
    rows = gdist.pair_row_indices(5, lo, hi)
    assert torch.equal(part["rep_tensor"], mol["rep_tensor"].reshape(25, 5, 5)[rows])


def test_payload_layout_and_exchange_selection_on_cpu():
    """Host logic of the in-library exchange: 16-byte aligned payload segments, and the fallbacks that keep CPU / gloo runs on
    the host framework's collective (no library exchange, no graph capture)."""
    from graddft_b200 import distributed as gdist

    offs, total = gdist.packed_layout([2 * 43 * 43, 1, 43 * 43, 2 * 43 * 43])
    assert offs == [0, 3698, 3700, 5550] and total == 9248 and all(o % 2 == 0 for o in offs)
    assert gdist.packed_layout([]) == ([], 0)
    assert gdist.fock_comm(128, "cpu") is None                      # CPU tensors: torch.distributed
    assert gdist.exchange_is_capturable("cpu") is False
    assert gdist.payload_segment(0, [8, 1], "cpu") is None            # no process group: nothing to write into
    x = [torch.arange(5, dtype=torch.float64), torch.ones(3, dtype=torch.float64)]
    out = gdist.allreduce_sum_packed(x)                               # world size 1: passed through untouched
    assert out[0] is x[0] and out[1] is x[1]


def test_packed_eri_cache_is_keyed_by_tensor_object_and_version():
    from graddft_b200 import ops

    cache = []
    a = torch.zeros(4, 4, 4, 4, dtype=torch.float64)
    e1 = ops._cache_entry(cache, a, ())
    assert ops._cache_entry(cache, a, ()) is e1
    a.add_(1.0)                                                       # in-place edit: a new entry
    e2 = ops._cache_entry(cache, a, ())
    assert e2 is not e1
    b = torch.zeros(4, 4, 4, 4, dtype=torch.float64)
    assert ops._cache_entry(cache, b, ()) is not e2 and ops._cache_entry(cache, b, (1, None)) is not ops._cache_entry(cache, b, ())
    del a
    import gc
    gc.collect()
    ops._cache_entry(cache, b, ())
    assert all(e["ref"]() is not None for e in cache)                 # entries of dead tensors are pruned
    assert ops.packed_eri_for(b) is None                              # CPU tensor: never packed
