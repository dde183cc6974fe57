"""Dense GEMM kernels (csrc/dense_gemm.cu) next to cuBLAS DGEMM at the DM21 trunk shape."""
import sys
import torch
sys.path.insert(0, ".")
from graddft_b200 import _lib, ops
from graddft_b200._lib import ptr, stream_ptr, wptr

dev = torch.device("cuda:0")
N, W = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000, 256
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64, device=dev)  # noqa: E731
x, k, kb, sc, bi, cot = rn(N, W), torch.eye(W, dtype=torch.float64, device=dev) + rn(W, W) / 16, rn(W), 1 + 0.1 * rn(W), 0.1 * rn(W), rn(N, W)
kt = k.t().contiguous()
L = _lib.lib()
flop = 2.0 * N * W * W


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def report(name, ms, flops=flop):
    print(f"{name:34s} {ms:8.3f} ms  {flops / ms / 1e9:6.2f} TFLOP/s", flush=True)


t_cublas = timeit(lambda: x @ k)
report("cuBLAS x @ k", t_cublas)
big = rn(8192, 8192)
report("cuBLAS 8192^3", timeit(lambda: big @ big), 2.0 * 8192 ** 3)
del big
out, xhat, rstd = torch.empty_like(x), torch.empty_like(x), torch.empty(N, dtype=torch.float64, device=dev)
report("gdft_dense_fwd (plain)", timeit(lambda: L.gdft_dense_fwd(stream_ptr(), N, W, W, ptr(x), ptr(kt), None, None, ptr(out))))
report("gdft_dense_fwd (+bias +res)", timeit(lambda: L.gdft_dense_fwd(stream_ptr(), N, W, W, ptr(x), ptr(kt), ptr(kb), ptr(x), ptr(out))))
report("gdft_dense_block_fwd (LN+ELU)", timeit(lambda: L.gdft_dense_block_fwd(stream_ptr(), N, W, ptr(x), ptr(kt), ptr(kb), ptr(sc), ptr(bi), 1e-6, ptr(out), ptr(xhat), ptr(rstd))))
ws = ops._dense_ws(N, W, W, dev)
zb, sb, bb, kbb = torch.empty_like(x), torch.empty(W, dtype=torch.float64, device=dev), torch.empty(W, dtype=torch.float64, device=dev), torch.empty(W, dtype=torch.float64, device=dev)
report("gdft_dense_block_bwd (chain)", timeit(lambda: L.gdft_dense_block_bwd(stream_ptr(), N, W, ptr(cot), ptr(k), ptr(out), ptr(xhat), ptr(rstd), ptr(sc), ptr(zb), ptr(sb), ptr(bb), ptr(kbb), wptr(ws), ws.numel())))
t_last = timeit(lambda: L.gdft_dense_block_bwd_last(stream_ptr(), N, W, ptr(cot), ptr(out), ptr(xhat), ptr(rstd), ptr(sc), ptr(zb), ptr(sb), ptr(bb), ptr(kbb), wptr(ws), ws.numel()))
print(f"{'gdft_dense_block_bwd_last':34s} {t_last:8.3f} ms  {32.0 * W * N / t_last / 1e6:6.0f} GB/s (32 W B/row)")
kbar = torch.empty(W, W, dtype=torch.float64, device=dev)
report("gdft_dense_bwd_weight (x^T z)", timeit(lambda: L.gdft_dense_bwd_weight(stream_ptr(), N, W, W, ptr(x), ptr(cot), ptr(kbar), wptr(ws), ws.numel())))
report("cuBLAS x.T @ cot", timeit(lambda: x.t() @ cot))
# the old path of one block: cuBLAS GEMM + K7 forward
y = x @ k
st = torch.empty(N, 2, dtype=torch.float64, device=dev)
t_k7 = timeit(lambda: L.gdft_dense_ln_elu_fwd(stream_ptr(), N, W, ptr(y), ptr(kb), ptr(x), ptr(sc), ptr(bi), 1e-6, ptr(out), ptr(st)))
print(f"old forward block: cuBLAS {t_cublas:.3f} + K7 {t_k7:.3f} = {t_cublas + t_k7:.3f} ms")
