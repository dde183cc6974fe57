"""Host memory copy rates on the GPU box: pageable -> pageable and pageable -> page-locked, 1 and 8 threads (ctypes.memmove
runs without the interpreter lock).  Explains the ceiling of the chi tail's staging copy."""
import ctypes, os, time
from concurrent.futures import ThreadPoolExecutor
import torch
n = 1 << 29  # 512 MiB
src = torch.ones(n // 8, dtype=torch.float64)
dst = torch.empty_like(src)
pin = torch.empty(n // 8, dtype=torch.float64).pin_memory()
def run(d, threads):
    step = n // threads
    def job(k): ctypes.memmove(d.data_ptr() + k * step, src.data_ptr() + k * step, step)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(job, range(threads)))
        t0 = time.perf_counter()
        for _ in range(3): list(ex.map(job, range(threads)))
        return 3 * n / (time.perf_counter() - t0) / 1e9
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for th in (1, 2, 4, 8):
    print(f"threads {th}: pageable->pageable {run(dst, th):6.1f} GB/s   pageable->pinned {run(pin, th):6.1f} GB/s")
g = torch.empty(n // 8, dtype=torch.float64, device="cuda")
for name, s in (("pinned", pin), ("pageable", src)):
    g.copy_(s); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3): g.copy_(s, non_blocking=True)
    torch.cuda.synchronize()
    print(f"H2D from {name}: {3 * n / (time.perf_counter() - t0) / 1e9:6.1f} GB/s")
