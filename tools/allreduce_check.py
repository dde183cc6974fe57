"""Multi-process check of the in-library exchange (run under torchrun, one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/allreduce_check.py
Both backends of distributed.FockComm (p2p: IPC-shared payloads + gdft_allreduce_fock_p2p; nccl: gdft_allreduce_fock) against
torch.distributed.all_reduce, bitwise agreement across ranks, latency per exchange, CUDA-graph capture of the p2p call, and the
sharded B3LYP predictor against the unsharded one."""
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import graddft_b200 as gd  # noqa: E402
from graddft_b200 import distributed as gdist  # noqa: E402
from graddft_b200.synthetic import synthetic_molecule  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True
for backend in ("p2p", "nccl"):
    for count in (2 * 43 * 43 + 1, 2 * 264 * 264 + 2, 2 * 400 * 400 + 2):
        comm = gdist.FockComm(count, dev, None, backend)
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        x = torch.randn(count, generator=g, dtype=torch.float64, device=dev)
        ref = x.clone()
        dist.all_reduce(ref)
        comm.buffer[:count].copy_(x)
        out = comm.allreduce(count).clone()
        err = float((out - ref).abs().max() / ref.abs().max())
        gathered = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(gathered, out)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        # latency: 50 exchanges back to back
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            comm.allreduce(count)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        e0.record()
        for _ in range(50):
            dist.all_reduce(ref)
        e1.record(); torch.cuda.synchronize()
        us_t = e0.elapsed_time(e1) * 1e3 / 50
        good = err < 1e-14 and same and comm.status() == 0
        ok &= good
        if rank == 0:
            print(f"{backend:5s} count={count:7d} rel err vs torch all_reduce {err:.1e}  bitwise equal across ranks {same}  {us:7.1f} us/exchange "
                  f"(torch.distributed {us_t:7.1f} us)  status {comm.status()}  {'ok' if good else 'FAIL'}", flush=True)
        if backend == "p2p" and count == 2 * 264 * 264 + 2:
            comm.buffer[:count].copy_(x)
            torch.cuda.synchronize(); dist.barrier()
            graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                with torch.cuda.graph(graph, stream=s):
                    comm.allreduce(count)
            comm.buffer[:count].copy_(x)
            torch.cuda.synchronize(); dist.barrier()
            graph.replay(); torch.cuda.synchronize()
            gerr = float((comm.buffer[:count] - out).abs().max())
            ok &= gerr == 0.0
            if rank == 0:
                print(f"p2p   CUDA-graph replay of the exchange: max abs diff vs eager {gerr:.1e}", flush=True)
        dist.barrier()
        comm.close()

# sharded predictor through the public API (payload written in place by the density VJP) vs the unsharded call
mol = synthetic_molecule(6000, 43, n_omega=1, seed=1984, mask_frac=0.0)
e_ref, f_ref = gd.energy_predictor(gd.B3LYP)(None, gd.molecule_from_tensors(mol, dev))
for backend in ("p2p", "nccl", "torch"):
    os.environ["GDFT_ALLREDUCE"] = backend
    for shard_eri in (False, True):
        m = gdist.shard_molecule(mol, rank, world, dev, None, shard_eri=shard_eri)
        e, f = gd.energy_predictor(gd.B3LYP)(None, m)
        de, df = abs(float(e) - float(e_ref)), float((f - f_ref).abs().max() / f_ref.abs().max())
        good = de < 1e-9 and df < 1e-9
        ok &= good
        if rank == 0:
            print(f"sharded B3LYP predictor [{backend}, shard_eri={shard_eri}]: |dE| {de:.1e}  rel dFock {df:.1e}  {'ok' if good else 'FAIL'}", flush=True)
flag = torch.tensor([0 if ok else 1], device=dev)
dist.all_reduce(flag)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 0 else 1)
