"""Kernel-level timing probe (development tool, not the bench): times each C-ABI kernel at the named shapes with
CUDA events and prints achieved FLOP/s / GB/s next to a cuBLAS DGEMM measurement taken in the same run."""
import sys, time, json, argparse
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops
from graddft_b200._lib import GDFT_RHO, GDFT_GRAD, GDFT_TAU, GDFT_LAPL, GDFT_HF
from graddft_b200.synthetic import synthetic_molecule

def timeit(fn, warm=2, rep=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(rep):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=2_000_000)
    ap.add_argument("--n", type=int, default=400)
    ap.add_argument("--eri_n", type=int, default=160)
    ap.add_argument("--W", type=int, default=0)
    ap.add_argument("--lapl", type=int, default=0)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    # cuBLAS DGEMM denominators
    for (M, Nn, K) in [(8192, 8192, 8192), (a.N // 4, 2 * a.n, a.n), (a.n, 2 * a.n, a.N // 4)]:
        A = torch.randn(M, K, dtype=torch.float64, device=dev)
        B = torch.randn(K, Nn, dtype=torch.float64, device=dev)
        ms = timeit(lambda: A @ B)
        print(f"cublas dgemm {M}x{Nn}x{K}: {ms:.3f} ms  {2*M*Nn*K/ms/1e9:.2f} TFLOP/s", flush=True)
        del A, B
    N, n = a.N, a.n
    mol = synthetic_molecule(N, n, n_omega=a.W, seed=1984, device=dev, with_eri=False, with_grad2=bool(a.lapl))
    basis = ops.PackedBasis(mol["ao"], mol["grad_ao"], mol.get("grad_n_ao2"), mol.get("chi"))
    D = mol["rdm1"]
    del mol["ao"], mol["grad_ao"]
    unit = 2.0 * N * n * n
    def rep(name, ms, flops=None, bytes_=None):
        s = f"{name:34s} {ms:9.3f} ms"
        if flops: s += f"  {flops/ms/1e9:7.2f} TFLOP/s"
        if bytes_: s += f"  {bytes_/ms/1e6:8.1f} GB/s"
        print(s, flush=True)
    ms = timeit(lambda: ops._density_fwd_raw(basis, D, GDFT_RHO)); rep("fwd RHO", ms, 2*unit)
    ms = timeit(lambda: ops._density_fwd_raw(basis, D, GDFT_RHO | GDFT_GRAD)); rep("fwd RHO|GRAD", ms, 2*unit)
    ms = timeit(lambda: ops._density_fwd_raw(basis, D, GDFT_RHO | GDFT_GRAD | GDFT_TAU)); rep("fwd RHO|GRAD|TAU", ms, 8*unit)
    rb = torch.randn(N, 2, dtype=torch.float64, device=dev); gb = torch.randn(N, 2, 3, dtype=torch.float64, device=dev)
    tb = torch.randn(N, 2, dtype=torch.float64, device=dev)
    ms = timeit(lambda: ops._density_bwd_raw(basis, GDFT_RHO, rb, None, None, None)); rep("bwd RHO", ms, 2*unit)
    ms = timeit(lambda: ops._density_bwd_raw(basis, GDFT_RHO | GDFT_GRAD, rb, gb, None, None)); rep("bwd RHO|GRAD", ms, 2*unit)
    ms = timeit(lambda: ops._density_bwd_raw(basis, GDFT_RHO | GDFT_GRAD | GDFT_TAU, rb, gb, tb, None)); rep("bwd RHO|GRAD|TAU", ms, 8*unit)
    rho, grho, tau, _, _ = ops._density_fwd_raw(basis, D, GDFT_RHO | GDFT_GRAD | GDFT_TAU)
    lapl = torch.randn(N, 2, dtype=torch.float64, device=dev) * rho
    for name in ["B88_SET", "B3LYP_SET", "DM21_INPUTS", "DM21_MGGA"]:
        F = ops.lib().gdft_pointwise_ncols(ops._lib.PW_IDS[name])
        ms = timeit(lambda: ops._Pointwise.apply(ops._lib.PW_IDS[name], 1e-30, rho, grho, tau, lapl)); rep(f"pw fwd {name}", ms, None, 8.0*N*(12+F))
        r2, g2 = rho.clone().requires_grad_(True), grho.clone().requires_grad_(True)
        out = ops.pointwise(name, r2, g2, tau, lapl)
        ob = torch.ones_like(out)
        ms = timeit(lambda: torch.autograd.grad(out, (r2, g2), ob, retain_graph=True)); rep(f"pw bwd {name}", ms, None, 8.0*N*(12+F+8))
    w = mol["weights"]; c = torch.ones(1, 2, dtype=torch.float64, device=dev); d = torch.randn(N, 2, dtype=torch.float64, device=dev)
    ms = timeit(lambda: ops.xc_integrate(c, d, w)); rep("integrate fwd F=2", ms, None, 8.0*N*3)
    del basis, rb, gb, tb, rho, grho, tau, lapl
    torch.cuda.empty_cache()
    ne = a.eri_n
    eri = torch.randn(ne, ne, ne, ne, dtype=torch.float64, device=dev); P = torch.randn(ne, ne, dtype=torch.float64, device=dev)
    ms = timeit(lambda: ops._eri_j_raw(P, eri)); rep(f"eri J n={ne}", ms, None, 8.0*ne**4)
    ms = timeit(lambda: ops._eri_jt_raw(P, eri)); rep(f"eri J^T n={ne}", ms, None, 8.0*ne**4)
    big = torch.empty(2**30, dtype=torch.float64, device=dev); big2 = torch.empty_like(big)
    ms = timeit(lambda: big2.copy_(big)); rep("torch copy 8 GiB", ms, None, 2*8.0*2**30)

main()
