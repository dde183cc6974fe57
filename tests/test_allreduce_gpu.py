"""The library's own exchange kernel (gdft_allreduce_fock_p2p, csrc/allreduce.cu) on ONE device: several ranks driven by one
process (gdft_comm_connect_local), each on its own stream, as they would be on the GPUs of one box.  Checks the sum, that
every rank ends with bitwise the same payload, fixed-order (run-to-run) reproducibility, repeated exchanges (epochs), odd
counts, and the zero-copy hand-over from the density VJP (ops.density_bwd_into).  The multi-process / multi-GPU path (IPC
handles, NCCL variant) is exercised by tools/allreduce_check.py under torchrun and by bench.py --gpus N."""
import ctypes

import pytest
import torch

from graddft_b200 import _lib, ops
from graddft_b200.distributed import _DevicePointer, packed_layout
from graddft_b200.synthetic import synthetic_molecule

pytestmark = pytest.mark.gpu
F64 = torch.float64


class LocalRanks:
    def __init__(self, world, capacity, dev):
        self.L = L = _lib.lib()
        self.world, self.dev = world, dev
        self.comms = []
        for r in range(world):
            c = ctypes.c_void_p()
            assert L.gdft_comm_create(r, world, capacity, ctypes.byref(c)) == 0
            self.comms.append(c)
        arr = (ctypes.c_void_p * world)(*[c.value for c in self.comms])
        for c in self.comms:
            assert L.gdft_comm_connect_local(c, arr) == 0
        self.cap = int(L.gdft_comm_capacity(self.comms[0]))
        self.bufs = [torch.as_tensor(_DevicePointer(int(L.gdft_comm_buffer(c)), self.cap), device=dev) for c in self.comms]
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(world)]

    def allreduce(self, count):
        cur = torch.cuda.current_stream()
        for r in range(self.world):
            self.streams[r].wait_stream(cur)
            with torch.cuda.stream(self.streams[r]):
                assert self.L.gdft_allreduce_fock_p2p(_lib.stream_ptr(), self.comms[r], count) == 0
        for s in self.streams:
            cur.wait_stream(s)

    def status(self):
        out = []
        for c in self.comms:
            st, ep = ctypes.c_int(0), ctypes.c_ulonglong(0)
            assert self.L.gdft_comm_status(c, ctypes.byref(st), ctypes.byref(ep)) == 0
            out.append((st.value, ep.value))
        return out

    def close(self):
        torch.cuda.synchronize()
        for c in self.comms:
            self.L.gdft_comm_destroy(c)


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_p2p_allreduce_sum_and_bitwise_agreement(cuda_device, world):
    count = 2 * 43 * 43 + 1  # odd on purpose
    R = LocalRanks(world, count, cuda_device)
    try:
        g = torch.Generator(device=cuda_device).manual_seed(world)
        for it in range(3):
            parts = [torch.randn(count, generator=g, dtype=F64, device=cuda_device) * (10.0 ** (r - 2)) for r in range(world)]
            for r in range(world):
                R.bufs[r][:count].copy_(parts[r])
            R.allreduce(count)
            torch.cuda.synchronize()
            want = parts[0].clone()
            for r in range(1, world):
                want += parts[r]  # rank order, the order of the kernel
            for r in range(world):
                assert torch.equal(R.bufs[r][:count], want), (world, it, r)
        assert R.status() == [(0, 3)] * world
    finally:
        R.close()


def test_p2p_allreduce_large_payload_and_argument_checks(cuda_device):
    count = 2 * 400 * 400 + 2
    R = LocalRanks(2, count, cuda_device)
    try:
        a = torch.randn(count, dtype=F64, device=cuda_device)
        b = torch.randn(count, dtype=F64, device=cuda_device)
        R.bufs[0][:count].copy_(a)
        R.bufs[1][:count].copy_(b)
        R.allreduce(count)
        torch.cuda.synchronize()
        assert torch.equal(R.bufs[0][:count], a + b) and torch.equal(R.bufs[1][:count], a + b)
        L = _lib.lib()
        assert L.gdft_allreduce_fock_p2p(_lib.stream_ptr(), R.comms[0], R.cap + 2) == 1  # beyond the payload
        assert L.gdft_allreduce_fock_p2p(_lib.stream_ptr(), None, 4) == 5
        assert L.gdft_comm_create(3, 2, 16, ctypes.byref(ctypes.c_void_p())) == 5
        assert L.gdft_allreduce_fock(None, _lib.stream_ptr(), None, 4) == 5
    finally:
        R.close()


def test_density_vjp_writes_the_payload_in_place(cuda_device):
    """Two 'ranks' hold the two halves of the grid; each VJP lands in its payload segment without a copy and the exchange
    reproduces the single-device V_xc (sum of the two partials, rank order)."""
    N, n = 4096, 37
    mol = synthetic_molecule(N, n, seed=1984, device=cuda_device, with_eri=False, mask_frac=0.0)
    offs, total = packed_layout([2 * n * n, 1])
    R = LocalRanks(2, total, cuda_device)
    try:
        g = torch.Generator(device=cuda_device).manual_seed(2)
        rb, gb = torch.randn(N, 2, generator=g, dtype=F64, device=cuda_device), torch.randn(N, 2, 3, generator=g, dtype=F64, device=cuda_device)
        halves = []
        for r, (lo, hi) in enumerate(((0, N // 2), (N // 2, N))):
            basis = ops.PackedBasis(mol["ao"][lo:hi].contiguous(), mol["grad_ao"][lo:hi].contiguous())
            with ops.density_bwd_into(R.bufs[r][offs[0]:offs[0] + 2 * n * n]):
                v = ops.density_transpose(basis, rb[lo:hi].contiguous(), gb[lo:hi].contiguous())
            assert v.data_ptr() == R.bufs[r].data_ptr()  # no copy: the reduce epilogue produced the payload
            halves.append(v.clone())
            R.bufs[r][offs[1]] = float(r + 1)
        R.allreduce(total)
        torch.cuda.synchronize()
        want = halves[0] + halves[1]
        for r in range(2):
            assert torch.equal(R.bufs[r][:2 * n * n].view(2, n, n), want)
            assert float(R.bufs[r][offs[1]]) == 3.0
        full = ops.density_transpose(ops.PackedBasis(mol["ao"], mol["grad_ao"]), rb, gb)
        assert float((want - full).abs().max() / full.abs().max()) < 1e-13
    finally:
        R.close()
