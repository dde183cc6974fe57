"""Timing of the fused SCF-tail kernels in isolation (development tool)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from graddft_b200 import ops, evaluate
dev = torch.device("cuda:0"); F64 = torch.float64
for n in (43, 64, 21):
    m = 10
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64, device=dev)
    X = rn(n, n); S = X @ X.T / n + torch.eye(n, dtype=F64, device=dev); L_inv = evaluate.overlap_factor(S)
    z = torch.zeros((m, 2, n, n), dtype=F64, device=dev); fv, ev, gram = z.clone(), z.clone(), torch.zeros((2, m, m), dtype=F64, device=dev)
    F, D = rn(2, n, n), rn(2, n, n)
    for c in range(12): ops.scf_diis_step(c, F + 0.1 * rn(2, n, n), D, S, L_inv, fv, ev, gram)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for c in range(12, 112): ops.scf_diis_step(c, F, D, S, L_inv, fv, ev, gram)
    b.record(); torch.cuda.synchronize()
    t_d = a.elapsed_time(b) * 10
    evals, V = torch.linalg.eigh(F + F.transpose(1, 2)); occ = torch.zeros(2, n, dtype=F64, device=dev); occ[:, :5] = 1
    ops.scf_occupy(evals, V, L_inv, occ); torch.cuda.synchronize()
    a.record()
    for _ in range(100): ops.scf_occupy(evals, V, L_inv, occ)
    b.record(); torch.cuda.synchronize()
    print(f"n={n}: scf_diis_step {t_d:.1f} us/call (back to back, includes launch), scf_occupy {a.elapsed_time(b) * 10:.1f} us/call", flush=True)
