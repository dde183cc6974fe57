"""Timing / sweep-count probe of the cluster eigensolver (gdft_sym_eigh_ex, n > 64) next to torch.linalg.eigh."""
import sys
import torch
sys.path.insert(0, ".")
from graddft_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


for n in (65, 90, 96, 128, 160, 200, 264, 320):
    A = torch.randn(2, n, n, generator=g, dtype=torch.float64)
    A = (A + A.transpose(1, 2)).to(dev)
    info = torch.zeros(2, dtype=torch.int32, device=dev)
    t_cold = timeit(lambda: ops.sym_eigh(A, info=info))
    sw_cold = info.tolist()
    w, V = ops.sym_eigh(A)
    line = f"n={n:4d} cold {t_cold:7.3f} ms sweeps {sw_cold} ({t_cold / max(sw_cold):.3f} ms/sweep)"
    for eps in (1e-3, 1e-5, 1e-7):
        P = torch.randn(2, n, n, generator=g, dtype=torch.float64).to(dev)
        A2 = A + eps * (P + P.transpose(1, 2))
        t = timeit(lambda: ops.sym_eigh(A2, V, info=info))
        line += f" | warm {eps:g}: {t:6.3f} ms {info.tolist()}"
    t_lib = timeit(lambda: torch.linalg.eigh(A))
    line += f" | torch.linalg.eigh {t_lib:6.3f} ms"
    if n <= 90:
        import os
        os.environ["GDFT_EIGH_ONE_CTA"] = "1"
        line += f" | one-CTA kernel {timeit(lambda: ops.sym_eigh(A)):6.3f} ms"
        os.environ["GDFT_EIGH_ONE_CTA"] = "0"
    if n >= 128:
        import os
        for cl in ("8", "16"):
            os.environ["GDFT_EIGH_CLUSTER"] = cl
            tc = timeit(lambda: ops.sym_eigh(A, info=info))
            P = torch.randn(2, n, n, generator=g, dtype=torch.float64).to(dev)
            A2 = A + 1e-5 * (P + P.transpose(1, 2))
            tw = timeit(lambda: ops.sym_eigh(A2, V, info=info))
            line += f" | cluster {cl}: cold {tc:6.3f} warm {tw:6.3f} {info.tolist()}"
        del os.environ["GDFT_EIGH_CLUSTER"]
    print(line, flush=True)
