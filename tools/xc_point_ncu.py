"""The one-pass per-point XC kernel (gdft_xc_point_fused) at the C4 grid size inside a profiler range, next to the chain of
kernels it replaces -- for `ncu --set full --profile-from-start off` and for CUDA-event timings.  python tools/xc_point_ncu.py"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops

dev = torch.device("cuda:0")
N = 2_000_000
g = torch.Generator(device=dev).manual_seed(7)
rho = torch.rand((N, 2), generator=g, dtype=torch.float64, device=dev) + 0.01
grho = torch.randn((N, 2, 3), generator=g, dtype=torch.float64, device=dev) * 0.3
lapl = torch.randn((N, 2), generator=g, dtype=torch.float64, device=dev)
ehf = -torch.rand((1, 2, N), generator=g, dtype=torch.float64, device=dev)
w = torch.rand((N,), generator=g, dtype=torch.float64, device=dev)
cases = {"B88_SET": ((1.0, 1.0), grho, None, None), "B3LYP_SET": ((0.8, 0.72, 0.19, 0.81, 0.2), grho, lapl, ehf)}


def timed(fn, reps=20):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def chain(name, coef, gr, la, eh):
    """what the generic first-order build launches for the same result: features, exact-exchange column, quadrature, their VJPs"""
    r, gg = rho.clone().requires_grad_(True), gr.clone().requires_grad_(True)
    ll = la.clone().requires_grad_(True) if la is not None else None
    with ops.first_order_build():
        d = ops.pointwise(name, r, gg, None, ll, clip=1e-30)
        if eh is not None:
            d = torch.cat([d, eh.sum(dim=(0, 1)).unsqueeze(1)], dim=1)
        c = torch.tensor([list(coef)], dtype=torch.float64, device=dev)
        e = ops.xc_integrate(c, d, w, 1e-30)
        torch.autograd.grad(e, [r, gg] + ([ll] if ll is not None else []))


with torch.no_grad():
    for name, (coef, gr, la, eh) in cases.items():
        fused = lambda: ops.xc_point_fused(name, 1e-30, coef, rho, gr, None, la, eh, w)
        t_f = timed(fused)
        with torch.enable_grad():
            t_c = timed(lambda: chain(name, coef, gr, la, eh))
        nin = 2 + 6 + (2 if la is not None else 0) + (2 if eh is not None else 0) + 1
        nout = 2 + 6 + (2 if la is not None else 0) + (2 if eh is not None else 0)
        gb = 8 * N * (nin + nout) / 1e9
        print(f"{name}: fused {t_f * 1e3:.1f} us = {gb / t_f:.2f} TB/s on {gb:.3f} GB algorithmic; generic chain {t_c * 1e3:.1f} us")
    torch.cuda.synchronize(); torch.cuda.profiler.start()
    for name, (coef, gr, la, eh) in cases.items():
        ops.xc_point_fused(name, 1e-30, coef, rho, gr, None, la, eh, w)
    torch.cuda.synchronize(); torch.cuda.profiler.stop()
