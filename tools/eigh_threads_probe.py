"""Development probe: the two spin blocks of a 264 x 264 eigenproblem on one stream vs two host threads + two streams."""
import sys, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 264
g = torch.Generator().manual_seed(1)
A = torch.randn(2, n, n, generator=g, dtype=torch.float64)
A = (A + A.transpose(1, 2)).to(dev)
pool = ThreadPoolExecutor(2)
streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

def batched():
    return torch.linalg.eigh(A)

def one(s):
    with torch.cuda.stream(streams[s]):
        return torch.linalg.eigh(A[s])

def threaded():
    cur = torch.cuda.current_stream()
    for st in streams:
        st.wait_stream(cur)
    futs = [pool.submit(one, s) for s in range(2)]
    res = [f.result() for f in futs]
    for st in streams:
        cur.wait_stream(st)
    return torch.stack([r[0] for r in res]), torch.stack([r[1] for r in res])

def wall(fn, rep=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(rep): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / rep * 1e3

w0, V0 = batched(); w1, V1 = threaded(); torch.cuda.synchronize()
print("same eigenvalues:", bool(torch.equal(w0, w1)), " max |dV|:", float((V0.abs() - V1.abs()).abs().max()))
print(f"n={n}: batched {wall(batched):.2f} ms   two threads/streams {wall(threaded):.2f} ms   single matrix {wall(lambda: torch.linalg.eigh(A[0])):.2f} ms")
