"""One launch each of the fused forward and reverse GEMM kernels at the DM21 trunk shape (for ncu)."""
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
out, xhat, rstd = torch.empty_like(x), torch.empty_like(x), torch.empty(N, dtype=torch.float64, device=dev)
ws = ops._dense_ws(N, W, W, dev)
zb, sb, bb, kbb = torch.empty_like(x), torch.empty(W, dtype=torch.float64, device=dev), torch.empty(W, dtype=torch.float64, device=dev), torch.empty(W, dtype=torch.float64, device=dev)
L.gdft_dense_fwd(stream_ptr(), N, W, W, ptr(x), ptr(kt), None, None, ptr(out))
L.gdft_dense_block_fwd(stream_ptr(), N, W, ptr(x), ptr(kt), ptr(kb), ptr(sc), ptr(bi), 1e-6, ptr(out), ptr(xhat), ptr(rstd))
L.gdft_dense_block_bwd(stream_ptr(), N, W, ptr(cot), ptr(k), ptr(out), ptr(xhat), ptr(rstd), ptr(sc), ptr(zb), ptr(sb), ptr(bb), ptr(kbb), wptr(ws), ws.numel())
torch.cuda.synchronize()
