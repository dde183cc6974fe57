import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import interface
dev = torch.device("cuda:0")
n, rows, chunk = 264, 4096, 1024
nu_host = torch.randn(rows, n, n, dtype=torch.float64)
t0 = time.perf_counter(); up = interface._Uploader(dev, chunk, n); torch.cuda.synchronize(); print("uploader ctor", round((time.perf_counter() - t0) * 1e3, 1), "ms")
for rep in range(2):
    for k in range(4):
        t0 = time.perf_counter()
        d, slot = up.stage(nu_host[k * chunk:(k + 1) * chunk])
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        up.consumed(slot)
        print(f"rep {rep} chunk {k}: stage() host time {1e3 * (t1 - t0):6.1f} ms, + wait for H2D {1e3 * (t2 - t1):6.1f} ms")
import ctypes
from graddft_b200.interface import _copy_rows, _stage_pool, _STAGE_THREADS
print("threads", _STAGE_THREADS)
src = nu_host[:chunk]; dst = up.pinned[0][:chunk]
for _ in range(3):
    t0 = time.perf_counter()
    m = chunk; parts = _STAGE_THREADS
    futs = [_stage_pool().submit(_copy_rows, dst, src, m * k // parts, m * (k + 1) // parts) for k in range(parts)]
    [f.result() for f in futs]
    dt = time.perf_counter() - t0
    print(f"pool copy of one chunk: {1e3 * dt:.1f} ms = {src.numel() * 8 / dt / 1e9:.1f} GB/s")
