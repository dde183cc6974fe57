"""Grid- and ERI-row-sharded predictor / SCF loop, world_size 2 (both ranks on cuda:0, gloo transport so that one
GPU suffices), against the unsharded result and the CPU oracle.  SURVEY.md section 8e."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
E_TOL, F_RTOL = 1e-8, 1e-7


def _worker(rank, world, port, q, shard_eri, N, n):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist
    from graddft_b200.synthetic import synthetic_molecule

    dev = torch.device("cuda:0")
    mol = synthetic_molecule(N, n, n_omega=2, seed=1984, mask_frac=0.0)
    m = gdist.shard_molecule(mol, rank, world, dev, shard_eri=shard_eri)
    e, f = gd.energy_predictor(gd.B3LYP)(None, m)
    out = gd.diff_scf_loop(gd.B3LYP, cycles=3)(None, m)
    torch.cuda.synchronize()
    q.put((rank, float(e), f.cpu(), float(out.energy), out.rdm1.cpu()))
    dist.barrier()
    dist.destroy_process_group()


# n = 96 is beyond the one-CTA Jacobi kernel: the two ranks split the spin blocks of the library eigensolve between them
@pytest.mark.parametrize("shard_eri,N,n", [(False, 3000, 20), (True, 3000, 20), (False, 1500, 96)])
def test_sharded_b3lyp_predictor_and_scf(cuda_device, shard_eri, N, n):
    import oracle
    import graddft_b200 as gd
    from graddft_b200.synthetic import synthetic_molecule

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, shard_eri, N, n)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda t: t[0])
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)

    mol = synthetic_molecule(N, n, n_omega=2, seed=1984, mask_frac=0.0)
    e_ref, f_ref = oracle.predict_b3lyp(mol)
    m = gd.molecule_from_tensors(mol, cuda_device)
    e1, f1 = gd.energy_predictor(gd.B3LYP)(None, m)
    scf1 = gd.diff_scf_loop(gd.B3LYP, cycles=3)(None, m)
    for rank, e, f, e_scf, rdm1 in res:
        assert abs(e - float(e_ref)) < E_TOL and abs(e - float(e1)) < E_TOL
        assert float((f - f_ref).abs().max() / f_ref.abs().max()) < F_RTOL
        assert abs(e_scf - float(scf1.energy)) < 1e-7
        assert float((rdm1 - scf1.rdm1.cpu()).abs().max()) < 1e-6
    # both ranks hold the same (replicated) Fock matrix bit for bit: the all-reduce result is identical everywhere
    assert torch.equal(res[0][2], res[1][2])
    assert torch.equal(res[0][4], res[1][4])  # ... and the same rdm1 after the SCF cycles (spin-split eigensolve included)
