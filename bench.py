#!/usr/bin/env python
"""bench.py -- XC build (forward + VJP) throughput on B200, next to the reference formulas on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[3], the synthetic XC sweep -- 2,000,000 grid points x 400 AOs,
GGA functional (LSDA + B88 exchange columns), float64.  A "step" is one XC build through the public API
(`graddft_b200.xc_energy_and_grads` = value_and_grad of Functional.xc_energy w.r.t. rdm1, grad_dft/train.py:86-121):
rho, grad rho -> per-point energy densities -> weighted grid integral E_xc, and the VJP back to V_xc[2,n,n].
With N > 1 GPUs the grid rows are sharded (strong scaling: the 2M-point grid is fixed) and one NCCL all-reduce of
[E_xc | V_xc] closes each build.  The packed basis (25.6 GB at N=1) is far larger than L2, so every step streams
from HBM ("l2": "inputs larger than L2").

`value`: builds/s with rdm1 resident in HBM.  `e2e`: the same call with rdm1 arriving from pinned host memory and
[E_xc | V_xc] returned to pinned host memory inside the timed region.  `roofline`: the dominant kernel
(density_bwd_kernel, the split-K aoT.M GEMM) against the FP64 GEMM rate of cuBLAS measured in this run.
`cpu_baseline` / `--impl reference`: the oracle's restatement of the reference einsums + autograd (torch-CPU,
float64, all host cores) on a bounded row sample of the same workload, scaled linearly in N.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

WORKLOADS = {
    "c4": dict(N=2_000_000, n=400, desc="synthetic XC sweep 2M grid pts x 400 AOs (BASELINE configs[3]), GGA (LSDA+B88), fwd+VJP"),
    "c3": dict(N=500_000, n=264, desc="benzene/def2-TZVP-shaped grid 500k pts x 264 AOs, GGA (LSDA+B88), fwd+VJP"),
}
CPU_SAMPLE_ROWS = 100_000


# ---------------------------------------------------------------------------------------------------------
def cpu_xc_build_rate(N_full: int, n: int, rows: int, steps: int, warmup: int):
    """Oracle (reference einsums + torch-CPU autograd) on `rows` grid rows; returns (builds/s scaled to N_full, cores, s/step)."""
    import oracle
    from graddft_b200.synthetic import synthetic_molecule

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mol = synthetic_molecule(rows, n, seed=1984, with_eri=False, with_grad2=False)

    def step():
        D = mol["rdm1"].clone().requires_grad_(True)
        e = oracle.xc_energy_of_rdm1(D, mol, "B88")
        (g,) = torch.autograd.grad(e, D)
        return float(e.detach()), g

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return (rows / N_full) / dt, cores, dt


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, cores, dt = cpu_xc_build_rate(wl["N"], wl["n"], CPU_SAMPLE_ROWS, max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)))
    sample = (f"{CPU_SAMPLE_ROWS} of {wl['N']} grid rows per step, n={wl['n']}; oracle einsums + torch-CPU autograd, float64; "
              f"builds/s scaled linearly in N (NumPy/torch-CPU restatement of the reference einsums, not JAX-CPU: jax is not installed)")
    line = {
        "impl": "reference", "metric": "xc_build_fwd_vjp_per_s", "value": rate, "unit": "builds/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "N": wl["N"], "n": wl["n"]},
        "cpu_baseline": {"value": rate, "unit": "builds/s", "cores": cores, "kind": "port", "sample": sample, "sample_s_per_step": dt},
        "e2e": {"value": rate, "unit": "builds/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])), mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measure_dgemm_tflops(dev, m=8192, reps=5):
    a = torch.randn(m, m, dtype=torch.float64, device=dev)
    b = torch.randn(m, m, dtype=torch.float64, device=dev)
    for _ in range(2):
        a @ b
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * m ** 3 / best / 1e9



# ---------------------------------------------------------------------------------------------------------
# second headline metric: jitted-SCF iterations/s (diff_scf_loop = make_jitted_scf_loop, grad_dft/evaluate.py:917)
# ---------------------------------------------------------------------------------------------------------
SCF_SHAPES = {
    # BASELINE.json configs[1]: H2O / def2-TZVP (n = 43), level-3 grid (~34k points), B3LYP
    "c2": dict(N=34_000, n=43, desc="H2O/def2-TZVP-shaped: 34k grid pts x 43 AOs, B3LYP (LSDA+B88+VWN+LYP+HF), DIIS SCF"),
    # benzene / def2-TZVP-shaped (configs[2] shape), B3LYP; rep_tensor 38.9 GB
    "c3": dict(N=500_000, n=264, desc="benzene/def2-TZVP-shaped: 500k grid pts x 264 AOs, B3LYP, DIIS SCF, rep_tensor 38.9 GB"),
    # BASELINE configs[2]: DM21 (11 -> 256 x 6 -> 3 network, two HF ranges) energy + Fock matrix, same shape
    "c3_dm21": dict(N=500_000, n=264, desc="benzene/def2-TZVP-shaped: 500k grid pts x 264 AOs, DM21 (seeded weights) energy_predictor call"),
}


def _scf_shard(N, n, rank, world, dev, n_omega=1):
    """Rank-local shard of a synthetic B3LYP-ready molecule: grid rows seeded per rank, n x n data replicated, the
    (p,q) rows of an 8-fold-symmetric PSD rep_tensor built directly as a row block (never materialised whole)."""
    from graddft_b200 import distributed as gdist
    from graddft_b200.synthetic import synthetic_molecule

    lo, hi = gdist.shard_bounds(N, rank, world)
    mol = synthetic_molecule(hi - lo, n, n_omega=n_omega, seed=1984 + rank, device=dev, with_eri=False, mask_frac=0.0)
    small = synthetic_molecule(8, n, seed=1984, device=dev, with_eri=False)
    for k in ("rdm1", "mo_coeff", "mo_occ", "mo_energy", "h1e", "s1e", "nuclear_repulsion"):
        mol[k] = small[k]
    mol["weights"] = mol["weights"] * ((hi - lo) / N)
    mol["omegas"] = [0.0, 0.4][:n_omega]
    g = torch.Generator(device=dev).manual_seed(4242)
    Q = 2 * n
    B = torch.randn(Q, n, n, generator=g, dtype=torch.float64, device=dev)
    B2 = (0.5 * (B + B.transpose(1, 2))).reshape(Q, n * n)
    r0, r1 = gdist.shard_bounds(n * n, rank, world, align=32) if world > 1 else (0, n * n)
    mol["rep_tensor"] = (B2[:, r0:r1].T @ B2).div_(Q).reshape(r1 - r0, n, n)
    del B, B2
    if world == 1:
        mol["rep_tensor"] = mol["rep_tensor"].reshape(n, n, n, n)
        return gdist_molecule(mol, dev, None)
    return gdist_molecule(mol, dev, gdist.GridShard(None, rank, world, r0))


def gdist_molecule(mol, dev, shard):
    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist

    m = gd.molecule_from_tensors(mol, dev)
    if shard is not None:
        gdist.attach_shard(m, shard)
    m.packed_basis
    return m


def scf_leg(shape_key, rank, world, dev, timed_ms, dgemm_tf, hbm_gbs):
    """ms per SCF iteration of diff_scf_loop(B3LYP): slope between a 2-cycle and a 6-cycle run (each iteration =
    DIIS extrapolation + generalised eigenproblem + occupations + rdm1 + one full Fock build)."""
    import graddft_b200 as gd
    from graddft_b200 import ops

    sh = SCF_SHAPES[shape_key]
    N, n = sh["N"], sh["n"]
    dm21 = shape_key.endswith("_dm21")
    m = _scf_shard(N, n, rank, world, dev, n_omega=2 if dm21 else 1)
    functional = gd.DM21() if dm21 else gd.B3LYP
    params = functional.generate_DM21_weights(device=dev) if dm21 else None
    if dm21:
        # BASELINE configs[2]: DM21 neural functional energy + gradient (one energy_predictor call = E and the Fock matrix)
        pred = gd.energy_predictor(functional)
        with torch.no_grad():
            for _ in range(2):
                e, f = pred(params, m)
            ms_pred = min(timed_ms(lambda: pred(params, m), 3) / 3.0 for _ in range(2))
        res = {"workload": sh["desc"], "N": N, "n": n, "predict_ms": ms_pred, "predicts_per_s": 1e3 / ms_pred,
               "energy_finite": bool(torch.isfinite(e)) and bool(torch.isfinite(f).all())}
        del m, e, f
        torch.cuda.empty_cache()
        return res
    # make_jitted_scf_loop = diff_scf_loop captured into a CUDA graph on first use (n <= 104; larger eigenproblems go
    # through cuSOLVER, whose status word forces the eager loop) -- the analogue of the reference's jax.jit
    loops = {c: gd.make_jitted_scf_loop(gd.B3LYP, cycles=c) for c in (2, 6)}
    eager = {c: gd.diff_scf_loop(gd.B3LYP, cycles=c) for c in (2, 6)}
    out = None
    with torch.no_grad():
        for c in (2, 6):
            out = loops[c](None, m)  # warm-up + capture (workspaces, cuSOLVER handles)
            eager[c](None, m)
        ms, ms_eager = {}, {}
        for c in (2, 6):
            ms[c] = min(timed_ms(lambda: loops[c](None, m), 1) for _ in range(3))
            ms_eager[c] = min(timed_ms(lambda: eager[c](None, m), 1) for _ in range(3))
        per_iter = (ms[6] - ms[2]) / 4.0
        ops.TIMING = {}
        out = eager[2](None, m)
        torch.cuda.synchronize()
        timing, ops.TIMING = ops.TIMING, None

    def avg(name):
        ev = timing.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))

    res = {"workload": sh["desc"], "N": N, "n": n, "iter_per_s": 1e3 / per_iter, "ms_per_iter": per_iter,
           "ms_loop_2_cycles": ms[2], "ms_loop_6_cycles": ms[6], "eager_ms_per_iter": (ms_eager[6] - ms_eager[2]) / 4.0,
           "cuda_graph": bool(n <= ops.lib().gdft_sym_eigh_max_n() and world == 1), "energy_finite": bool(torch.isfinite(out.energy))}
    if rank == 0:
        rows = m.rep_tensor.shape[0] * (m.rep_tensor.shape[1] if m.rep_tensor.dim() == 4 else 1)
        eri_ms = avg("gdft_eri_jk")
        eri_bytes = 8.0 * rows * n * n
        res["kernels_ms"] = {"density_fwd": avg("gdft_density_fwd"), "density_bwd": avg("gdft_density_bwd"), "eri_j": eri_ms}
        if eri_bytes > 2.5e8:  # larger than L2: a DRAM figure
            res["roofline_eri"] = {"bound": "hbm", "kernel": "eri_j_kernel (rep_tensor (pq)x(rt) sweep)", "achieved": eri_bytes / eri_ms / 1e6,
                                   "peak": hbm_gbs[0], "unit": "GB/s", "frac": eri_bytes / eri_ms / 1e6 / hbm_gbs[0],
                                   "bytes_per_launch": eri_bytes, "ms_per_launch": eri_ms, "peak_source": hbm_gbs[1]}
        # roofline of one iteration: B3LYP Fock build = 16 GEMM units (rho, grad, lapl fwd + VJP) + 2 (HF Fock) at the
        # measured DGEMM rate, plus one rep_tensor sweep at the HBM peak (per-GPU shares)
        unit = 2.0 * (N / world) * n * n
        t_roof = 18.0 * unit / (dgemm_tf * 1e9) + eri_bytes / (hbm_gbs[0] * 1e6)
        res["roofline_iter"] = {"ms_at_roofline": t_roof, "frac": t_roof / per_iter,
                                "model": "18 units x 2*N*n^2 FLOP at measured cuBLAS DGEMM + 8*rows*n^2 B at HBM peak"}
    del m, out
    torch.cuda.empty_cache()
    return res


def train_leg(rank, world, dev, timed_ms):
    """BASELINE configs[4]: one training step (non-SCF energy loss, grad_dft/train.py:312-359,480-535) of the DM21-shaped
    neural functional on a batch of 64 small synthetic molecules (n_i = 12 + floor(88 U), N_i = 1e4 (1 + 3 U), seed
    1993), molecules sharded over the ranks (balanced by N n^2), one gradient all-reduce per step, Adam update."""
    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist
    from graddft_b200.synthetic import synthetic_molecule

    g = torch.Generator().manual_seed(1993)
    shapes = [(int(1e4 * (1 + 3 * float(torch.rand((), generator=g)))), 12 + int(88 * float(torch.rand((), generator=g)))) for _ in range(64)]
    mine = gdist.shard_molecules([N * n * n for N, n in shapes], rank, world)
    mols = {i: gd.molecule_from_tensors(synthetic_molecule(shapes[i][0], shapes[i][1], n_omega=2, seed=1993 + i, device=dev, mask_frac=0.0), dev)
            for i in mine}
    for m in mols.values():
        m.packed_basis
    fun = gd.DM21()
    params = {k: v.requires_grad_(True) for k, v in fun.generate_DM21_weights(device=dev).items()}
    leaves = list(params.values())
    opt = torch.optim.Adam(leaves, lr=1e-4)
    predictor = gd.non_scf_predictor(fun)
    truths = {i: torch.tensor(-1.0 - 0.01 * i, dtype=torch.float64, device=dev) for i in mine}

    def step():
        total = torch.zeros((), dtype=torch.float64, device=dev)
        # what mse_energy_loss evaluates (the loss reads .energy only): features per molecule, one network pass per group
        energies = predictor.energy_only_batch(params, [mols[i] for i in mine])
        for i, energy in zip(mine, energies):
            total = total + ((energy - truths[i]) / mols[i].mo_occ.sum()) ** 2
        loss = total / 64
        grads = torch.autograd.grad(loss, leaves)
        grads, loss = gdist.allreduce_gradients(list(grads), loss.detach())
        for p_, g_ in zip(leaves, grads):
            p_.grad = g_
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(2):
        loss = step()
    ms = min(timed_ms(step, 1) for _ in range(3))
    return {"workload": "64 synthetic molecules (n 12..100, N 1e4..4e4), DM21-shaped functional (11 -> 256 x 6 -> 3), non-SCF "
                        "energy loss + Adam step; molecules sharded over ranks, one gradient all-reduce per step",
            "ms_per_step": ms, "molecules_per_s": 64e3 / ms, "molecules_on_rank0": len(mine), "loss_finite": bool(torch.isfinite(loss))}


def cpu_scf_iter_rate(N, n):
    """Oracle SCF iteration (B3LYP) on the host cores: slope between 1- and 3-cycle loops."""
    import oracle
    from graddft_b200.synthetic import synthetic_molecule

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    mol = synthetic_molecule(N, n, n_omega=1, seed=1984, mask_frac=0.0)
    t = {}
    for c in (1, 3):
        t0 = time.perf_counter()
        oracle.diff_scf_loop_energy(mol, oracle.predict_b3lyp, c)
        t[c] = time.perf_counter() - t0
    per_iter = (t[3] - t[1]) / 2.0
    return 1.0 / per_iter, cores, per_iter


def chi_leg(rank, world, dev, timed_ms, hbm_gbs):
    """Row f4: the chi-generation tail at the benzene shape (n = 264): `rows` grid points per GPU with their nu[r] (n x n
    per point, 557 KB) resident in HBM, streamed once by gdft_chi_contract; roofline = 8 n^2 bytes per point against the
    HBM peak.  `e2e` (1 GPU only): the same points with nu arriving from host memory in 1024-point chunks through the
    double-buffered uploader of generate_chi_tensor (PCIe-bound by construction: nu is produced on the host by libcint)."""
    import torch.distributed as dist
    from graddft_b200 import interface, ops
    n, rows, chunk = 264, 9472, 1024  # 148 SMs x 2 CTAs x 8 points x 4 groups
    g = torch.Generator(device=dev).manual_seed(1984 + rank)
    ao = torch.randn(rows, n, generator=g, dtype=torch.float64, device=dev)
    D = torch.randn(2, n, n, generator=g, dtype=torch.float64, device=dev)
    nu = torch.randn(rows, n, n, generator=g, dtype=torch.float64, device=dev)  # 4.6 GB >> L2
    coords = torch.arange(rows, dtype=torch.float64, device=dev)[:, None].expand(rows, 3)
    chi = torch.empty((rows, 1, 2, n), dtype=torch.float64, device=dev)

    def step():
        ops.chi_contract_(chi, 0, 0, ao, D, nu)

    for _ in range(3):
        step()
    steps = 10
    ms = timed_ms(step, steps) / steps
    bytes_per_launch = 8.0 * rows * n * n
    hbm_gbs, hbm_src = hbm_gbs
    out = {"workload": f"chi tail at the benzene shape: {rows} grid points per GPU x nu[264,264] per point (one omega), nu resident in HBM",
           "rows_per_gpu": rows, "n": n, "ms_per_launch": ms, "points_per_s": world * rows / (ms / 1e3),
           "roofline": {"bound": "hbm", "kernel": "chi_contract_kernel (one pass over nu)", "achieved": bytes_per_launch / ms / 1e6,
                        "peak": hbm_gbs, "unit": "GB/s", "frac": bytes_per_launch / ms / 1e6 / hbm_gbs, "bytes_per_launch": bytes_per_launch,
                        "peak_source": hbm_src},
           "finite": bool(torch.isfinite(chi).all())}
    if world == 1:
        host_rows = 4096
        nu_host = nu[:host_rows].cpu()
        cidx = torch.arange(host_rows, dtype=torch.float64)[:, None].expand(host_rows, 3)

        def e2e():
            interface.generate_chi_tensor(D, ao[:host_rows], cidx, lambda c, omega: nu_host[int(c[0, 0]):int(c[0, 0]) + len(c)], [0.0], chunk)

        e2e()
        ms_e = timed_ms(e2e, 3) / 3
        out["e2e"] = {"value": host_rows / (ms_e / 1e3), "unit": "points/s", "h2d_bytes_per_step": 8 * host_rows * n * n,
                      "d2h_bytes_per_step": 0, "ms_per_step": ms_e, "h2d_GBps": 8e-6 * host_rows * n * n / ms_e,
                      "note": "nu chunks of 1024 points from pageable host memory through two pinned buffers"}
        nu_pinned = nu_host.pin_memory()

        def e2e_pinned():
            interface.generate_chi_tensor(D, ao[:host_rows], cidx, lambda c, omega: nu_pinned[int(c[0, 0]):int(c[0, 0]) + len(c)], [0.0], chunk)

        e2e_pinned()
        ms_p = timed_ms(e2e_pinned, 3) / 3
        out["e2e_pinned_source"] = {"value": host_rows / (ms_p / 1e3), "unit": "points/s", "ms_per_step": ms_p,
                                    "h2d_GBps": 8e-6 * host_rows * n * n / ms_p,
                                    "note": "the provider hands over page-locked chunks: no staging copy, H2D overlapped with the kernel"}
    return out


def hbm_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs ('of measured')"
        except (KeyError, ValueError):
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s ('of fallback': MEASURED_PEAKS.json absent)"


def run_ours(args, wl):
    import torch.distributed as dist

    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist
    from graddft_b200 import ops
    from graddft_b200.synthetic import synthetic_molecule

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: graddft_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if ops.lib().gdft_device_supported() != 1:
        raise SystemExit("libgdft_b200 targets sm_100a (B200) only")

    N, n = wl["N"], wl["n"]
    lo, hi = gdist.shard_bounds(N, rank, world)
    Nloc = hi - lo
    # rank-local rows of the synthetic grid (seeded per rank); rdm1 & co from rank 0's seed, replicated
    mol = synthetic_molecule(Nloc, n, seed=1984 + rank, device=dev, with_eri=False, with_grad2=False)
    small = synthetic_molecule(8, n, seed=1984, device=dev, with_eri=False, with_grad2=False)
    for k in ("rdm1", "mo_coeff", "mo_occ", "mo_energy", "h1e", "s1e"):
        mol[k] = small[k]
    molecule = gd.molecule_from_tensors(mol, dev)
    molecule.packed_basis  # pack once: ao / grad_ao are constant across SCF iterations and training steps
    del mol
    torch.cuda.empty_cache()

    functional = gd.B88
    rdm1_dev = molecule.rdm1.clone()
    rdm1_host = rdm1_dev.cpu().pin_memory()
    payload = torch.empty(1 + 2 * n * n, dtype=torch.float64, device=dev)
    out_host = torch.empty(1 + 2 * n * n, dtype=torch.float64).pin_memory()

    def build(rdm1):
        exc, vxc, _ = gd.xc_energy_and_grads(functional, None, rdm1, molecule, create_graph=False)
        if world > 1:
            exc, vxc = gdist.allreduce_xc(exc, vxc, buf=payload)
        return exc, vxc

    def step_resident():
        return build(rdm1_dev)

    def step_e2e():
        rdm1_dev.copy_(rdm1_host, non_blocking=True)
        exc, vxc = build(rdm1_dev)
        gdist.pack_xc(exc, vxc, payload) if world == 1 else None
        out_host.copy_(payload, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    dgemm_tf = measure_dgemm_tflops(dev) if rank == 0 else None

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.lib().gdft_launch_count()
    ops.TIMING = {}
    profile_range = os.environ.get("GDFT_BENCH_PROFILE_RANGE") == "1"  # ncu --profile-from-start off: only the timed steps
    if profile_range:
        torch.cuda.profiler.start()
    ms_total = timed(step_resident, args.steps)
    if profile_range:
        torch.cuda.profiler.stop()
    timing, ops.TIMING = ops.TIMING, None
    launches = (ops.lib().gdft_launch_count() - launches0) / args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()

    packed_gb = 4 * Nloc * molecule.packed_basis.npad * 8 / 1e9
    scf = None
    if not args.no_scf:
        del molecule
        torch.cuda.empty_cache()
        if world > 1:
            dgemm_all = torch.tensor([dgemm_tf or 0.0], dtype=torch.float64, device=dev)
            dist.broadcast(dgemm_all, 0)
            dgemm_tf_all = float(dgemm_all)
        else:
            dgemm_tf_all = dgemm_tf
        scf = {}
        for key in (("c2", "c3", "c3_dm21") if world == 1 else ("c3", "c3_dm21")):
            try:
                scf[key] = scf_leg(key, rank, world, dev, timed, dgemm_tf_all, hbm_peak())
            except Exception as exc:  # the headline XC line must survive a failure of the secondary leg
                scf[key] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            scf["c5_training"] = train_leg(rank, world, dev, timed)
        except Exception as exc:
            scf["c5_training"] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            torch.cuda.empty_cache()
            scf["chi_tail"] = chi_leg(rank, world, dev, timed, hbm_peak())
        except Exception as exc:
            scf["chi_tail"] = {"error": f"{type(exc).__name__}: {exc}"}

    # sanity: the result that went to the host is finite
    assert bool(torch.isfinite(out_host).all()), "non-finite XC build"

    if rank == 0:
        def avg_ms(name):
            ev = timing.get(name, [])
            return sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))

        bwd_ms, fwd_ms = avg_ms("gdft_density_bwd"), avg_ms("gdft_density_fwd")
        flop_half = 4.0 * Nloc * n * n  # 2 GEMM units per call (both spins): 2 * (2 N n^2)
        ach_bwd = flop_half / bwd_ms / 1e9
        ach_fwd = flop_half / fwd_ms / 1e9
        traffic = None
        tpath = ROOT / "profiles" / "r1_traffic.json"
        if world == 1 and tpath.exists():  # dram bytes per launch from the committed ncu --set full capture of this workload
            traffic = json.loads(tpath.read_text()).get(args.workload, {}).get("density_bwd_kernel")
        value = args.steps / (ms_total / 1e3)
        e2e = args.steps / (ms_e2e / 1e3)
        cpu_rate, cores, cpu_dt = (None, None, None)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu_rate, cores, cpu_dt = cpu_xc_build_rate(N, n, CPU_SAMPLE_ROWS, 3, 1)
            cpu = {"value": cpu_rate, "unit": "builds/s", "cores": cores, "kind": "port",
                   "sample": f"{CPU_SAMPLE_ROWS} of {N} grid rows, n={n}, 3 steps after 1 warm-up ({cpu_dt:.2f} s/step), scaled linearly in N; "
                             "oracle restatement of the reference einsums + torch-CPU autograd (not JAX-CPU: jax is not installed)"}
        line = {
            "metric": "xc_build_fwd_vjp_per_s", "value": value, "unit": "builds/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "N": N, "n": n, "rows_per_gpu": Nloc, "functional": "B88 (LSDA+B88 columns)",
                       "parallelism": f"grid-sharded x{world}, one all-reduce of [E_xc|V_xc] per build" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (packed basis %.1f GB per GPU)" % packed_gb},
            "e2e": {"value": e2e, "unit": "builds/s", "h2d_bytes_per_step": rdm1_host.numel() * 8, "d2h_bytes_per_step": out_host.numel() * 8,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "density_bwd_kernel (gdft_density_bwd: aoT.M split-K DMMA GEMM)",
                         "achieved": ach_bwd, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach_bwd / dgemm_tf, "traffic": traffic,
                         "flop_per_launch": flop_half, "ms_per_launch": bwd_ms,
                         "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json carries no FP64 figure); 'of measured'"},
            "roofline_fwd": {"bound": "tensor", "kernel": "density_fwd_kernel (gdft_density_fwd: ao.D DMMA GEMM + fused row dots)",
                             "achieved": ach_fwd, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach_fwd / dgemm_tf,
                             "flop_per_launch": flop_half, "ms_per_launch": fwd_ms},
            "xc_build_tflops": 8.0 * N * n * n / (ms_total / args.steps) / 1e9,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if scf is not None:
            if world == 1 and not args.no_cpu_baseline and "error" not in scf.get("c2", {"error": 1}):
                rate, cores_s, dt = cpu_scf_iter_rate(SCF_SHAPES["c2"]["N"], SCF_SHAPES["c2"]["n"])
                scf["c2"]["cpu_baseline"] = {"value": rate, "unit": "iter/s", "cores": cores_s, "kind": "port",
                                             "sample": f"full H2O-shaped molecule, oracle diff_scf_loop (torch-CPU float64), {dt * 1e3:.0f} ms/iter"}
            line["scf"] = {"metric": "jitted_scf_iter_per_s (diff_scf_loop = make_jitted_scf_loop)", "unit": "iter/s", **scf}
        emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(text: str) -> None:
    """The ONE JSON line goes to the process's real stdout; everything else that libraries print there (NCCL's
    version banner, for one) has been diverted to stderr by `main`."""
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=list(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-scf", dest="no_scf", action="store_true", help="skip the secondary SCF-iteration leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
