#!/usr/bin/env python
"""bench.py -- the two headline metrics of BASELINE.json on B200, next to the reference formulas on the host CPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload WL] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workloads (`--workload`, named in config.workload); each prints ONE JSON line with its own metric / roofline /
cpu_baseline / e2e (and, for N > 1, a `parity` object: the all-reduced result against a single-GPU recompute on rank 0):

  c4        (default) BASELINE.json configs[3]: synthetic XC sweep, 2,000,000 grid points x 400 AOs, GGA functional
            (LSDA + B88 exchange columns), float64.  A step is one XC build through the public API
            (`graddft_b200.xc_energy_and_grads` = value_and_grad of Functional.xc_energy w.r.t. rdm1,
            grad_dft/train.py:86-121): rho, grad rho -> per-point energy densities -> E_xc, and the VJP back to
            V_xc[2,n,n].  8 N n^2 FLOP per build.  With N > 1 GPUs the grid rows are sharded (strong scaling) and one
            all-reduce of [E_xc | V_xc] closes each build.
  c4_mgga   the same sweep with a meta-GGA functional (the `densities(..., "MGGA")` feature library, rho / grad rho / tau):
            32 N n^2 FLOP per build.
  c3        the GGA build at the benzene/def2-TZVP shape (500k x 264).
  scf_c2    jitted SCF iterations/s, H2O/def2-TZVP-shaped (34k x 43), B3LYP, `make_jitted_scf_loop` (configs[1]).
  scf_c3    jitted SCF iterations/s, benzene/def2-TZVP-shaped (500k x 264, rep_tensor 38.9 GB), B3LYP.
  dm21_c3   DM21 `energy_predictor` calls/s (energy + Fock matrix) at the benzene shape (configs[2]).
  train_c5  training steps/s on a batch of 64 small molecules, DM21-shaped functional, molecule-sharded (configs[4]).

The default run (c4) also carries the other legs under the key "scf" (skip with --no-extras).  All inputs are far
larger than L2 or are re-streamed from HBM every step ("l2" in config).  `value`: inputs resident in HBM.  `e2e`: the
same call with its per-step inputs arriving from pinned host memory and its result returned to pinned host memory inside
the timed region.  `roofline`: the dominant kernel against cuBLAS DGEMM measured in this run (FP64 tensor bound; both
the 8192^3 burst figure and the tall-skinny shape are printed) or against the measured HBM peak.
`cpu_baseline` / `--impl reference`: the oracle's restatement of the reference einsums + autograd (torch-CPU, float64,
all host cores) on a bounded sample of the same workload; the line states the sample, the steps it really ran and the
wall time of each.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

XC_SHAPES = {
    "c4": dict(N=2_000_000, n=400, kind="gga", units=4,
               desc="synthetic XC sweep 2M grid pts x 400 AOs (BASELINE configs[3]), GGA (LSDA+B88), fwd+VJP"),
    "c4_mgga": dict(N=2_000_000, n=400, kind="mgga", units=16,
                    desc="synthetic XC sweep 2M grid pts x 400 AOs (BASELINE configs[3]), meta-GGA (rho, grad rho, tau feature library), fwd+VJP"),
    "c3": dict(N=500_000, n=264, kind="gga", units=4, desc="benzene/def2-TZVP-shaped grid 500k pts x 264 AOs, GGA (LSDA+B88), fwd+VJP"),
}
SCF_SHAPES = {
    # BASELINE.json configs[1]: H2O / def2-TZVP (n = 43), level-3 grid (~34k points), B3LYP
    "c2": dict(N=34_000, n=43, desc="H2O/def2-TZVP-shaped: 34k grid pts x 43 AOs, B3LYP (LSDA+B88+VWN+LYP+HF), DIIS SCF (make_jitted_scf_loop)"),
    # benzene / def2-TZVP-shaped (configs[2] shape), B3LYP; rep_tensor 38.9 GB
    "c3": dict(N=500_000, n=264, desc="benzene/def2-TZVP-shaped: 500k grid pts x 264 AOs, B3LYP, DIIS SCF (make_jitted_scf_loop), rep_tensor 38.9 GB"),
    # BASELINE configs[2]: DM21 (11 -> 256 x 6 -> 3 network, two HF ranges) energy + Fock matrix, same shape
    "c3_dm21": dict(N=500_000, n=264, desc="benzene/def2-TZVP-shaped: 500k grid pts x 264 AOs, DM21 (seeded weights) energy_predictor call"),
}
WORKLOADS = list(XC_SHAPES) + ["scf_c2", "scf_c3", "dm21_c3", "train_c5"]
TRAIN_DESC = ("64 synthetic molecules (n 12..100, N 1e4..4e4), DM21-shaped functional (11 -> 256 x 6 -> 3), non-SCF "
              "energy loss + Adam step; molecules sharded over ranks, one gradient all-reduce per step (BASELINE configs[4])")
REF_BUDGET_S = 150.0  # wall-time target of a whole `--impl reference` run of an XC workload
E_TOL, V_RTOL = 1e-8, 1e-7  # BASELINE.json: total energy within 1e-8 Ha, V_xc and gradients within 1e-7 relative


# =========================================================================================================
# CPU side: the oracle (reference einsums + torch-CPU autograd) on bounded samples.  Only this section and the
# parity checkers at the bottom of it touch `oracle/`.
# =========================================================================================================
def _cpu_threads() -> int:
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return cores


def _oracle_xc(mol, D, kind):
    import oracle

    if kind == "gga":
        return oracle.xc_energy_of_rdm1(D, mol, "B88")
    rho = oracle.density(D, mol["ao"])
    grho = oracle.grad_density(D, mol["ao"], mol["grad_ao"])
    tau = oracle.kinetic_density(D, mol["grad_ao"])
    d = oracle.abs_clip(oracle.mgga_feature_densities(rho, grho, tau, "MGGA"))
    return oracle.xc_energy(torch.ones(1, d.shape[1], dtype=d.dtype), d, mol["weights"])


def cpu_xc_build(N_full: int, n: int, kind: str, rows: int, steps: int, warmup: int):
    """`steps` timed oracle XC builds (forward + autograd VJP) on `rows` grid rows after `warmup` untimed ones.
    Returns dict(rate scaled linearly to N_full rows, cores, seconds per executed step, steps executed)."""
    from graddft_b200.synthetic import synthetic_molecule

    cores = _cpu_threads()
    mol = synthetic_molecule(rows, n, seed=1984, with_eri=False, with_grad2=False)

    def step():
        D = mol["rdm1"].clone().requires_grad_(True)
        e = _oracle_xc(mol, D, kind)
        (g,) = torch.autograd.grad(e, D)
        return float(e.detach()), g

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    return {"rate": (rows / N_full) / dt, "cores": cores, "s_per_step": dt, "steps": steps, "rows": rows}


def cpu_scf_iteration(N: int, n: int, grid_rows: int, eri_p: int, reps: int = 1):
    """Seconds per oracle `diff_scf_loop` iteration (B3LYP) at an (N, n) shape, composed from bounded samples that each
    scale linearly: the grid part of one predictor call (value_and_grad of E_xc, the feature re-evaluation and the
    explicit exact-exchange Fock term, grad_dft/train.py:147-213) on `grid_rows` rows; one rep_tensor sweep
    (molecule.py:811; the reference's second sweep inside nonXC is merged by XLA's CSE under the loop's jit and is not
    charged) on `eri_p` of the n leading (p) slabs; and the n x n part (CDIIS, generalised eigenproblem, occupations,
    rdm1) in full.  For the H2O shape everything is run in full (grid_rows = N, eri_p = n)."""
    import oracle
    from graddft_b200.synthetic import synthetic_molecule

    cores = _cpu_threads()
    mol = synthetic_molecule(grid_rows, n, n_omega=1, seed=1984, mask_frac=0.0, with_eri=False)
    g = torch.Generator().manual_seed(4242)
    Q = 2 * n
    B = torch.randn(Q, n, n, generator=g, dtype=torch.float64)
    B2 = (0.5 * (B + B.transpose(1, 2))).reshape(Q, n * n)
    eri_block = ((B2[:, : eri_p * n].T @ B2) / Q).reshape(eri_p, n, n, n)
    del B, B2
    D = mol["rdm1"]
    P = D.sum(0)

    def grid_part():
        rdm1 = D.detach().clone().requires_grad_(True)
        exc = oracle.xc_energy_of_rdm1(rdm1, mol, "B3LYP")
        (fxc,) = torch.autograd.grad(exc, rdm1)
        ao, gao = mol["ao"], mol["grad_ao"]
        feats = oracle.b3lyp_exhf_densities(oracle.density(D, ao), oracle.grad_density(D, ao, gao),
                                            oracle.lapl_density(D, ao, gao, mol["grad_n_ao2"]))
        chi = mol["chi"][:, :1]
        ehf = oracle.HF_energy_density(D, ao, chi).detach().requires_grad_(True)
        E = oracle.xc_energy(torch.tensor([oracle.B3LYP_COEFFS], dtype=ao.dtype), oracle.b3lyp_combine(feats, ehf), mol["weights"])
        (gg,) = torch.autograd.grad(E, ehf)
        return fxc, oracle.HF_fock(chi, gg, ao)

    def eri_part():
        return oracle.coulomb_potential(P, eri_block)

    fock = mol["h1e"].expand(2, n, n).contiguous() + 0.1 * D
    z = torch.zeros(10, 2, n, n, dtype=torch.float64)
    filled = torch.randn(10, 2, n, n, generator=g, dtype=torch.float64)  # ring buffers as after a few cycles (a zero Gram matrix is singular)
    data = (z.clone(), filled.clone(), torch.zeros(10, dtype=torch.float64), filled - filled.transpose(2, 3))
    nelecs = mol["mo_occ"].sum(dim=1).round().to(torch.int64)

    def small_part():
        F, _ = oracle.jittable_diis_run(mol["s1e"], D, fock, torch.tensor(-1.0, dtype=torch.float64), data, 3)
        e, c = oracle.safe_fock_solver(F, mol["s1e"])
        return oracle.make_rdm1(c, oracle.get_occ(e, nelecs))

    def best(fn):
        fn()
        ts = []
        for _ in range(max(1, reps)):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    tg, te, ts_ = best(grid_part), best(eri_part), best(small_part)
    per_iter = tg * (N / grid_rows) + te * (n / eri_p) + ts_
    return {"s_per_iter": per_iter, "cores": cores, "grid_s": tg, "eri_s": te, "small_s": ts_, "grid_rows": grid_rows, "eri_p": eri_p}


def cpu_dm21_predict(N: int, n: int, grid_rows: int, eri_p: int):
    """Seconds per oracle DM21 `energy_predictor` call at an (N, n) shape from bounded samples: the grid part on
    `grid_rows` rows (three network passes forward, two backward, as the reference evaluates them), one rep_tensor sweep
    on `eri_p` leading slabs."""
    import oracle
    from graddft_b200.synthetic import synthetic_molecule

    cores = _cpu_threads()
    mol = synthetic_molecule(grid_rows, n, n_omega=2, seed=1984, mask_frac=0.0, with_eri=False)
    g = torch.Generator().manual_seed(4242)
    Q = 2 * n
    B = torch.randn(Q, n, n, generator=g, dtype=torch.float64)
    B2 = (0.5 * (B + B.transpose(1, 2))).reshape(Q, n * n)
    eri_block = ((B2[:, : eri_p * n].T @ B2) / Q).reshape(eri_p, n, n, n)
    del B, B2
    params = oracle.dm21_mlp_init(seed=1984)
    # predict_dm21 sweeps a full [n,n,n,n] tensor; give it an n-independent stand-in of one slab and time the sweep apart
    small = dict(mol)
    P = mol["rdm1"].sum(0)

    def grid_part():
        rdm1 = mol["rdm1"].detach().clone().requires_grad_(True)
        exc = oracle.xc_energy_of_rdm1(rdm1, small, "DM21", params=params)
        (fxc,) = torch.autograd.grad(exc, rdm1)
        D, ao, gao, chi, w = mol["rdm1"], mol["ao"], mol["grad_ao"], mol["chi"], mol["weights"]
        rho, grho, tau = oracle.density(D, ao), oracle.grad_density(D, ao, gao), oracle.kinetic_density(D, gao)
        gd_, gc_ = oracle.dm21_densities(rho, grho, tau, "LDA"), oracle.dm21_coefficient_inputs(rho, grho, tau)
        ehf0 = oracle.HF_energy_density(D, ao, chi).detach()
        dens, cin = oracle.dm21_combine_densities(gd_, ehf0), oracle.dm21_combine_cinputs(gc_, ehf0)
        ehf = ehf0.clone().requires_grad_(True)
        (g1,) = torch.autograd.grad(oracle.xc_energy(oracle.dm21_mlp(params, cin), oracle.dm21_combine_densities(gd_, ehf), w), ehf)
        v1 = oracle.HF_fock(chi, g1, ao)
        ehf = ehf0.clone().requires_grad_(True)
        (g2,) = torch.autograd.grad(oracle.xc_energy(oracle.dm21_mlp(params, oracle.dm21_combine_cinputs(gc_, ehf)), dens, w), ehf)
        return fxc, v1, oracle.HF_fock(chi, g2, ao)

    grid_part()
    t0 = time.perf_counter()
    grid_part()
    tg = time.perf_counter() - t0
    oracle.coulomb_potential(P, eri_block)
    t0 = time.perf_counter()
    oracle.coulomb_potential(P, eri_block)
    te = time.perf_counter() - t0
    return {"s_per_call": tg * (N / grid_rows) + te * (n / eri_p), "cores": cores, "grid_s": tg, "eri_s": te, "grid_rows": grid_rows, "eri_p": eri_p}


def cpu_train_step(shapes, sample):
    """Seconds per oracle training step over the 64-molecule batch, from the molecules whose indices are in `sample`
    (forward energy + parameter gradient each), scaled by total N n^2 work."""
    import oracle
    from graddft_b200.synthetic import synthetic_molecule

    cores = _cpu_threads()
    params = {k: v.clone().requires_grad_(True) for k, v in oracle.dm21_mlp_init(seed=1984).items()}
    total = 0.0
    t_all = 0.0
    work = lambda N, n: N * (n * n + 8.0e5 / 8)  # density GEMM units + the network's ~8e5 FLOP per point  # noqa: E731
    for i in sample:
        N, n = shapes[i]
        mol = synthetic_molecule(N, n, n_omega=2, seed=1993 + i, mask_frac=0.0)
        t0 = time.perf_counter()
        e = oracle.xc_energy_of_rdm1(mol["rdm1"], mol, "DM21", params=params) + oracle.nonXC(mol["rdm1"].sum(0), mol["h1e"], mol["rep_tensor"], mol["nuclear_repulsion"])
        torch.autograd.grad((e + 1.0) ** 2, list(params.values()))
        t_all += time.perf_counter() - t0
        total += work(N, n)
    full = sum(work(N, n) for N, n in shapes)
    return {"s_per_step": t_all * full / total, "cores": cores, "sample_s": t_all, "sample": list(sample)}


def oracle_block_parity(molecule, functional_kind: str, rows: int = 4096):
    """Checker (rank 0): [E_xc | V_xc] of the kernels on the first `rows` grid rows of this rank's molecule against the
    CPU oracle on the same tensors."""
    import graddft_b200 as gd

    sub = {"ao": molecule.ao[:rows], "grad_ao": molecule.grad_ao[:rows], "weights": molecule.grid.weights[:rows], "rdm1": molecule.rdm1}
    host = {k: v.detach().cpu() for k, v in sub.items()}
    D = host["rdm1"].clone().requires_grad_(True)
    e_ref = _oracle_xc(host, D, functional_kind)
    (v_ref,) = torch.autograd.grad(e_ref, D)
    full = dict(sub)
    for k in ("mo_coeff", "mo_occ", "mo_energy", "h1e", "s1e"):
        full[k] = getattr(molecule, k)
    m = gd.molecule_from_tensors(full, molecule.rdm1.device)
    exc, vxc, _ = gd.xc_energy_and_grads(_xc_functional(functional_kind), None, m.rdm1, m, create_graph=False)
    dE = abs(float(exc) - float(e_ref))
    dV = float((vxc.cpu() - v_ref).abs().max() / v_ref.abs().max())
    return {"against": f"CPU oracle on the first {rows} grid rows", "dE": dE, "dV_rel": dV, "ok": bool(dE < E_TOL and dV < V_RTOL)}


# =========================================================================================================
def emit(text: str) -> None:
    """The ONE JSON line goes to the process's real stdout; everything else that libraries print there (NCCL's
    version banner, for one) has been diverted to stderr by `main`."""
    if _REAL_STDOUT is None:
        print(text, flush=True)
    else:
        os.write(_REAL_STDOUT, (text + "\n").encode())


_REAL_STDOUT = None


def _xc_config(key: str, world: int):
    """config of an XC workload: identical for both arms (the reference arm runs on "your arm's config")."""
    from graddft_b200 import distributed as gdist

    wl = XC_SHAPES[key]
    N, n = wl["N"], wl["n"]
    lo, hi = gdist.shard_bounds(N, 0, world)
    npad = (n + 7) // 8 * 8
    packed_gb = 4 * (hi - lo) * npad * 8 / 1e9
    return {"workload": wl["desc"], "N": N, "n": n, "rows_per_gpu": hi - lo,
            "functional": "B88 (LSDA+B88 columns)" if wl["kind"] == "gga" else "densities(MGGA) feature library, unit coefficients",
            "flop_per_build": 2.0 * wl["units"] * N * n * n,
            "parallelism": f"grid-sharded x{world}, one all-reduce of [E_xc|V_xc] per build" if world > 1 else "single GPU",
            "l2": "inputs larger than L2 (packed basis %.1f GB per GPU, streamed from HBM every step)" % packed_gb}


def _scf_config(key: str, world: int):
    sh = SCF_SHAPES[key]
    return {"workload": sh["desc"], "N": sh["N"], "n": sh["n"],
            "parallelism": (f"grid rows and rep_tensor (p,q) rows sharded x{world}, one all-reduce of [E_xc|V_xc|J|V_HF] per Fock build"
                            if world > 1 else "single GPU"),
            "l2": "inputs larger than L2 (packed basis + chi + rep_tensor streamed from HBM every iteration)" if sh["n"] > 100
                  else "26 MB rep_tensor + 47 MB packed basis per iteration: L2-resident by construction at this shape (latency-bound workload)"}


def _train_config(world: int):
    return {"workload": TRAIN_DESC, "molecules": 64, "parallelism": f"molecule-sharded x{world}, one gradient all-reduce per step" if world > 1 else "single GPU",
            "l2": "inputs larger than L2 (64 packed bases, ~9 GB, streamed every step)"}


def run_reference(args):
    """--impl reference: the oracle port (cpu_baseline.kind = "port": the reference itself needs jax/flax/pyscf, see
    DESIGN.md section 8) on the host cores, on a bounded sample of the same workload.  `steps` / `warmup` /
    `ms_per_step` are what was really executed; `value` is scaled to the full workload as `cpu_baseline.sample` says."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    key = args.workload
    note = "NumPy/torch-CPU restatement of the reference einsums (oracle/), not JAX-CPU: jax is not installed"
    if key in XC_SHAPES:
        wl = XC_SHAPES[key]
        # rows per step: as many as keep the whole run (every one of the --steps + --warmup steps really executed) within
        # REF_BUDGET_S seconds on this box, between 100k and 500k of the workload's rows (a calibration step on 20k rows)
        if "GDFT_REF_ROWS" in os.environ:
            rows = int(os.environ["GDFT_REF_ROWS"])
        else:
            cal = cpu_xc_build(wl["N"], wl["n"], wl["kind"], 20_000, 1, 1)
            per_row = cal["s_per_step"] / 20_000
            rows = int(REF_BUDGET_S / (per_row * (max(1, args.steps) + max(0, args.warmup))))
            rows = max(100_000, min(500_000, rows // 10_000 * 10_000))
        rows = min(rows, wl["N"])
        r = cpu_xc_build(wl["N"], wl["n"], wl["kind"], rows, max(1, args.steps), max(0, args.warmup))
        frac = rows / wl["N"]
        line = {"metric": "xc_build_fwd_vjp_per_s", "value": r["rate"], "unit": "builds/s", "steps": r["steps"], "warmup": max(0, args.warmup),
                "ms_per_step": 1e3 * r["s_per_step"], "units_per_step": frac, "scaling": "strong", "config": _xc_config(key, world),
                "sample": f"every step = oracle forward + autograd VJP on {rows} of {wl['N']} grid rows ({frac:.3f} of one build, "
                          f"{r['s_per_step']:.2f} s wall each), n = {wl['n']}; builds/s = {frac:.3f} / s_per_step (linear in rows); {note}"}
    elif key in ("scf_c2", "scf_c3"):
        sh = SCF_SHAPES[key[4:]]
        full = key == "scf_c2"
        r = cpu_scf_iteration(sh["N"], sh["n"], sh["N"] if full else 25_000, sh["n"] if full else 16, reps=max(1, min(args.steps, 5)))
        line = {"metric": "jitted_scf_iter_per_s", "value": 1.0 / r["s_per_iter"], "unit": "iter/s", "steps": max(1, min(args.steps, 5)), "warmup": 1,
                "ms_per_step": 1e3 * (r["grid_s"] + r["eri_s"] + r["small_s"]), "scaling": "strong", "config": _scf_config(key[4:], world),
                "sample": f"one oracle diff_scf_loop iteration composed of: grid part on {r['grid_rows']} of {sh['N']} rows ({r['grid_s']:.3f} s), one "
                          f"rep_tensor sweep on {r['eri_p']} of {sh['n']} leading slabs ({r['eri_s']:.3f} s), n x n part in full ({r['small_s']:.3f} s); "
                          f"each scaled linearly -> {r['s_per_iter']:.3f} s/iter; {note}"}
    elif key == "dm21_c3":
        sh = SCF_SHAPES["c3_dm21"]
        r = cpu_dm21_predict(sh["N"], sh["n"], 10_000, 16)
        line = {"metric": "dm21_energy_predictor_calls_per_s", "value": 1.0 / r["s_per_call"], "unit": "calls/s", "steps": 1, "warmup": 1,
                "ms_per_step": 1e3 * (r["grid_s"] + r["eri_s"]), "scaling": "strong", "config": _scf_config("c3_dm21", world),
                "sample": f"grid part on {r['grid_rows']} of {sh['N']} rows ({r['grid_s']:.2f} s) + rep_tensor sweep on {r['eri_p']} of {sh['n']} slabs "
                          f"({r['eri_s']:.3f} s), scaled linearly -> {r['s_per_call']:.1f} s/call; {note}"}
    else:
        shapes = _train_shapes()
        r = cpu_train_step(shapes, [0, 21, 42, 63])
        line = {"metric": "training_steps_per_s", "value": 1.0 / r["s_per_step"], "unit": "steps/s", "steps": 1, "warmup": 0,
                "ms_per_step": 1e3 * r["sample_s"], "scaling": "strong", "config": _train_config(world),
                "sample": f"molecules {r['sample']} of 64 (energy + parameter gradient each, {r['sample_s']:.2f} s), scaled by N (n^2 + 1e5) -> "
                          f"{r['s_per_step']:.1f} s/step; {note}"}
    sample = line.pop("sample")
    line.update({"impl": "reference", "n_gpus": args.gpus, "higher_is_better": True, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                 "cpu_baseline": {"value": line["value"], "unit": line["unit"], "cores": os.cpu_count() or 1, "kind": "port", "sample": sample},
                 "e2e": {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(json.dumps(line))


# =========================================================================================================
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


def measure_dgemm_tflops(dev, shape=(8192, 8192, 8192), reps=5):
    """cuBLAS DGEMM rate (best of `reps`) for an m x k times k x n product: the FP64 tensor-pipe denominator of the
    rooflines (MEASURED_PEAKS.json carries no FP64 figure)."""
    m, k, n = shape
    a = torch.randn(m, k, dtype=torch.float64, device=dev)
    b = torch.randn(k, n, dtype=torch.float64, device=dev)
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
    return 2.0 * m * k * n / best / 1e9


def hbm_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs ('of measured')"
        except (KeyError, ValueError):
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s ('of fallback': MEASURED_PEAKS.json absent)"


class Ctx:
    """Process-group plumbing and the timing rule of the contract: barrier + synchronize on both sides, CUDA events on
    the launching stream, MAX over ranks."""

    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl=ours) needs a CUDA device: graddft_b200 has no CPU path")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        from graddft_b200 import ops

        if ops.lib().gdft_device_supported() != 1:
            raise SystemExit("libgdft_b200 targets sm_100a (B200) only")
        self._dgemm = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, finish=None):
        """`finish` (optional) is called after the last step and before the end event: a step that leaves copies running
        on side streams makes the timing stream wait for them there, so they are inside the timed region."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def dgemm(self):
        """(8192^3 burst TFLOP/s, tall-skinny 500000 x 800 x 400 TFLOP/s), measured on rank 0 and broadcast."""
        if self._dgemm is None:
            v = torch.zeros(2, dtype=torch.float64, device=self.dev)
            if self.rank == 0:
                v[0] = measure_dgemm_tflops(self.dev)
                v[1] = measure_dgemm_tflops(self.dev, (500_000, 400, 800))
            if self.world > 1:
                self.dist.broadcast(v, 0)
            self._dgemm = (float(v[0]), float(v[1]))
        return self._dgemm

    def fail(self, what, parity):
        sys.stderr.write(f"bench.py: PARITY FAILURE in {what}: {json.dumps(parity)}\n")
        sys.stderr.flush()
        raise SystemExit(3)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def _xc_functional(kind: str):
    import graddft_b200 as gd

    if kind == "gga":
        return gd.B88
    fn = _xc_functional.__dict__.get("mgga")
    if fn is None:
        from graddft_b200.popular_functionals import _ConstRow

        fn = gd.Functional(coefficients=lambda self, *_: _ConstRow.get((1.0,) * 16),
                           energy_densities=lambda m, clip_cte=1e-30, *a, **k: gd.densities(m, "MGGA", clip_cte), needs=("rho", "grad", "tau"))
        _xc_functional.mgga = fn
    return fn


def _xc_shard(ctx, key, r):
    """Rank r's molecule of an XC workload: its rows of the synthetic grid (seeded per rank), rdm1 & co from seed 1984."""
    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist
    from graddft_b200.synthetic import synthetic_molecule

    wl = XC_SHAPES[key]
    lo, hi = gdist.shard_bounds(wl["N"], r, ctx.world)
    mol = synthetic_molecule(hi - lo, wl["n"], seed=1984 + r, device=ctx.dev, with_eri=False, with_grad2=False)
    small = synthetic_molecule(8, wl["n"], seed=1984, device=ctx.dev, with_eri=False, with_grad2=False)
    for k in ("rdm1", "mo_coeff", "mo_occ", "mo_energy", "h1e", "s1e"):
        mol[k] = small[k]
    m = gd.molecule_from_tensors(mol, ctx.dev)
    m.packed_basis  # pack once: ao / grad_ao are constant across SCF iterations and training steps
    return m


def run_xc(ctx, args, key, extras):
    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist
    from graddft_b200 import ops

    wl = XC_SHAPES[key]
    N, n, kind = wl["N"], wl["n"], wl["kind"]
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    molecule = _xc_shard(ctx, key, rank)
    Nloc = molecule.grid_size
    torch.cuda.empty_cache()
    functional = _xc_functional(kind)
    rdm1_dev = molecule.rdm1.clone()
    rdm1_host = rdm1_dev.cpu().pin_memory()
    # payload [V_xc(2,n,n) | E_xc | pad]: with several ranks it is the exchange buffer of the library's own all-reduce
    # (gdft_allreduce_fock_p2p over NVLink peer memory); the density VJP's second-stage reduce writes V_xc straight into it
    n2 = n * n
    if world > 1:
        comm = gdist.fock_comm(2 * n2 + 2, dev)
        payload = comm.buffer[:2 * n2 + 2] if comm is not None else torch.zeros(2 * n2 + 2, dtype=torch.float64, device=dev)
    else:
        comm = None
        payload = torch.zeros(2 * n2 + 2, dtype=torch.float64, device=dev)
    out_host = torch.empty(2 * n2 + 2, dtype=torch.float64).pin_memory()

    def build(rdm1):
        with ops.density_bwd_into(payload[:2 * n2]):
            exc, vxc, _ = gd.xc_energy_and_grads(functional, None, rdm1, molecule, create_graph=False)
        if world > 1:
            exc, vxc = gdist.allreduce_xc(exc, vxc)
            if comm is None:
                payload[:2 * n2].copy_(vxc.reshape(-1))
                payload[2 * n2:2 * n2 + 1].copy_(exc.reshape(1))
        else:
            if vxc.data_ptr() != payload.data_ptr():
                payload[:2 * n2].copy_(vxc.reshape(-1))
            payload[2 * n2:2 * n2 + 1].copy_(exc.reshape(1))
        return exc, vxc

    def step_resident():
        return build(rdm1_dev)

    # end to end: every step uploads its density matrix from pinned host memory and reads [V_xc | E_xc] back.  The copies run
    # on two copy streams with double-buffered device staging, so the upload of step k+1 and the read-back of step k overlap
    # the kernels of their neighbours (2 x 2.56 MB per step: at 8 GPUs the serialised copies were 3 % of the step)
    h2d_stream, d2h_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    rd = [rdm1_dev, rdm1_dev.clone()]
    stage = [torch.empty_like(payload), torch.empty_like(payload)]
    rd_free, stage_free = [None, None], [None, None]
    e2e_k = [0]

    def step_e2e():
        i = e2e_k[0] & 1
        e2e_k[0] += 1
        cur = torch.cuda.current_stream(dev)
        if rd_free[i] is not None:
            h2d_stream.wait_event(rd_free[i])      # the kernels of step k-2 have finished reading this buffer
        with torch.cuda.stream(h2d_stream):
            rd[i].copy_(rdm1_host, non_blocking=True)
            up = torch.cuda.Event()
            up.record()
        cur.wait_event(up)
        build(rd[i])
        rd_free[i] = torch.cuda.Event()
        rd_free[i].record()
        if stage_free[i] is not None:
            cur.wait_event(stage_free[i])          # the read-back of step k-2 has left this staging buffer
        stage[i].copy_(payload)                     # the payload itself is overwritten by the next build
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(d2h_stream):
            d2h_stream.wait_event(done)
            out_host.copy_(stage[i], non_blocking=True)
            stage_free[i] = torch.cuda.Event()
            stage_free[i].record()

    def finish_e2e():
        cur = torch.cuda.current_stream(dev)
        cur.wait_stream(d2h_stream)
        cur.wait_stream(h2d_stream)

    dgemm_tf, dgemm_ts = ctx.dgemm()
    warm = max(3, args.warmup)
    for _ in range(warm):
        step_resident()
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.lib().gdft_launch_count()
    ops.TIMING = {}
    profile_range = os.environ.get("GDFT_BENCH_PROFILE_RANGE") == "1"  # ncu --profile-from-start off: only the timed steps
    if profile_range:
        torch.cuda.profiler.start()
    ms_total = ctx.timed(step_resident, args.steps)
    if profile_range:
        torch.cuda.profiler.stop()
    timing, ops.TIMING = ops.TIMING, None
    launches = (ops.lib().gdft_launch_count() - launches0) / args.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = ctx.timed(step_e2e, args.steps, finish=finish_e2e)
    clocks = sampler.stop() if rank == 0 else None
    torch.cuda.synchronize()
    assert bool(torch.isfinite(out_host).all()), "non-finite XC build"

    # ---- value checks (outside the timed region) --------------------------------------------------------
    parity = None
    if world > 1:
        # the all-reduced [E_xc | V_xc] against a single-GPU recompute: rank 0 regenerates every rank's shard in turn,
        # runs the unsharded call on it and sums the partial results in rank order
        step_e2e()
        torch.cuda.synchronize()
        sharded = out_host.clone()
        del molecule
        torch.cuda.empty_cache()
        if rank == 0:
            acc = torch.zeros(2 * n2 + 2, dtype=torch.float64, device=dev)
            for r in range(world):
                m_r = _xc_shard(ctx, key, r)
                exc, vxc, _ = gd.xc_energy_and_grads(functional, None, m_r.rdm1, m_r, create_graph=False)
                acc[:2 * n2] += vxc.reshape(-1)
                acc[2 * n2] += exc
                del m_r, exc, vxc
                torch.cuda.empty_cache()
            acc = acc.cpu()
            dE = abs(float(acc[2 * n2] - sharded[2 * n2]))
            dV = float((acc[:2 * n2] - sharded[:2 * n2]).abs().max() / acc[:2 * n2].abs().max())
            parity = {"against": f"rank-0 single-GPU recompute of all {world} shards (same kernels, no collective)", "dE_vs_1gpu": dE, "dV_rel": dV,
                      "E_xc": float(sharded[2 * n2]), "exchange": (comm.backend if comm is not None else "torch.distributed"),
                      "comm_status": (comm.status() if comm is not None else None), "tolerance": {"dE": E_TOL, "dV_rel": V_RTOL}, "ok": bool(dE < E_TOL and dV < V_RTOL)}
        ctx.barrier()
        if parity is not None and not parity["ok"]:
            ctx.fail(f"{key} x{world}", parity)
    elif not args.no_cpu_baseline:
        parity = oracle_block_parity(molecule, kind)
        parity["tolerance"] = {"dE": E_TOL, "dV_rel": V_RTOL}
        if not parity["ok"]:
            ctx.fail(key, parity)
        del molecule
    else:
        del molecule
    torch.cuda.empty_cache()

    scf = extras(ctx, args) if extras is not None else None
    if rank != 0:
        return

    def avg_ms(name):
        ev = timing.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))

    bwd_ms, fwd_ms = avg_ms("gdft_density_bwd"), avg_ms("gdft_density_fwd")
    flop_half = wl["units"] / 2.0 * 2.0 * Nloc * n * n  # half of the build's GEMM units in each of K1 / K2
    ach_bwd, ach_fwd = flop_half / bwd_ms / 1e9, flop_half / fwd_ms / 1e9
    traffic = None
    tpath = ROOT / "profiles" / "r2_traffic.json"
    if not tpath.exists():
        tpath = ROOT / "profiles" / "r1_traffic.json"
    if world == 1 and tpath.exists():  # dram bytes per launch from the committed ncu --set full capture of this workload
        traffic = json.loads(tpath.read_text()).get(key, {}).get("density_bwd_kernel")
    value = args.steps / (ms_total / 1e3)
    e2e = args.steps / (ms_e2e / 1e3)
    line = {
        "metric": "xc_build_fwd_vjp_per_s", "value": value, "unit": "builds/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": _xc_config(key, world),
        "e2e": {"value": e2e, "unit": "builds/s", "h2d_bytes_per_step": rdm1_host.numel() * 8, "d2h_bytes_per_step": out_host.numel() * 8,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "density_bwd_kernel (gdft_density_bwd: aoT.M split-K DMMA GEMM)",
                     "achieved": ach_bwd, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach_bwd / dgemm_tf, "traffic": traffic,
                     "flop_per_launch": flop_half, "ms_per_launch": bwd_ms, "peak_tall_skinny": dgemm_ts, "frac_of_tall_skinny": ach_bwd / dgemm_ts,
                     "peak_source": "cuBLAS DGEMM measured in this run: 8192^3 burst (`peak`) and 500000x400x800 (`peak_tall_skinny`); "
                                    "MEASURED_PEAKS.json carries no FP64 figure; 'of measured'"},
        "roofline_fwd": {"bound": "tensor", "kernel": "density_fwd_kernel (gdft_density_fwd: ao.D DMMA GEMM + fused row dots)",
                         "achieved": ach_fwd, "peak": dgemm_tf, "unit": "TFLOP/s", "frac": ach_fwd / dgemm_tf,
                         "flop_per_launch": flop_half, "ms_per_launch": fwd_ms},
        "xc_build_tflops": 2.0 * wl["units"] * N * n * n / (ms_total / args.steps) / 1e9,
    }
    if parity is not None:
        line["parity"] = parity
    if world == 1 and not args.no_cpu_baseline:
        rows = 250_000 if kind == "gga" else 100_000
        r = cpu_xc_build(N, n, kind, min(rows, N), 3, 1)
        line["cpu_baseline"] = {"value": r["rate"], "unit": "builds/s", "cores": r["cores"], "kind": "port",
                                "sample": f"{r['rows']} of {N} grid rows, n={n}, {r['steps']} steps after 1 warm-up ({r['s_per_step']:.2f} s/step), scaled "
                                          "linearly in N; oracle restatement of the reference einsums + torch-CPU autograd (not JAX-CPU: jax is not installed)"}
    if scf is not None:
        line["scf"] = {"metric": "jitted_scf_iter_per_s (diff_scf_loop = make_jitted_scf_loop)", "unit": "iter/s", **scf}
    emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# second headline metric: jitted-SCF iterations/s (diff_scf_loop = make_jitted_scf_loop, grad_dft/evaluate.py:917)
# ---------------------------------------------------------------------------------------------------------
def _scf_tensors(N, n, rank, world, dev, n_omega=1, eri_rows=None):
    """Rank-local tensors of a synthetic B3LYP/DM21-ready molecule: grid rows seeded per rank, n x n data replicated, and a
    row block of an 8-fold-symmetric PSD rep_tensor built directly (never materialised whole unless asked for): with several
    ranks the rank's share of the (p >= q) pair rows, else the (p,q) rows `eri_rows` (default: all)."""
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
    if eri_rows is None and world > 1:
        # this rank's share of the n(n+1)/2 (p >= q) pair rows (balanced; the other half of the rows is their mirror image)
        p0, p1 = gdist.pair_bounds(n, rank, world)
        idx = gdist.pair_row_indices(n, p0, p1, dev)
        mol["rep_tensor"] = (B2[:, idx].T @ B2).div_(Q).reshape(p1 - p0, n, n)
        del B, B2
        return mol, p0
    r0, r1 = eri_rows if eri_rows is not None else (0, n * n)
    mol["rep_tensor"] = (B2[:, r0:r1].T @ B2).div_(Q).reshape(r1 - r0, n, n)
    del B, B2
    return mol, r0


def _scf_shard(N, n, rank, world, dev, n_omega=1):
    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist

    mol, r0 = _scf_tensors(N, n, rank, world, dev, n_omega)
    if world == 1:
        mol["rep_tensor"] = mol["rep_tensor"].reshape(n, n, n, n)
    m = gd.molecule_from_tensors(mol, dev)
    if world > 1:
        gdist.attach_shard(m, gdist.GridShard(None, rank, world, None, r0))
    m.packed_basis
    return m


def _scf_unsharded(N, n, world, dev, n_omega=1):
    """The WHOLE molecule of a sharded SCF workload on one GPU (rank 0's parity recompute): every rank's grid rows
    regenerated from its seed and concatenated, the full rep_tensor."""
    import graddft_b200 as gd

    parts = [_scf_tensors(N, n, r, world, dev, n_omega, eri_rows=(0, 0))[0] for r in range(world)]
    mol = dict(parts[0])
    for k in ("ao", "grad_ao", "grad_n_ao2", "chi", "weights", "coords"):
        if parts[0].get(k) is not None:
            mol[k] = torch.cat([p[k] for p in parts], dim=0)
    del parts
    mol["rep_tensor"] = _scf_tensors(8, n, 0, 1, dev, 0, eri_rows=(0, n * n))[0]["rep_tensor"].reshape(n, n, n, n)
    m = gd.molecule_from_tensors(mol, dev)
    m.packed_basis
    return m


def scf_leg(ctx, shape_key, steps=4, with_parity=True, with_e2e=False):
    """ms per SCF iteration of make_jitted_scf_loop(B3LYP): slope between a 2-cycle and a (2 + steps)-cycle call (each
    iteration = DIIS extrapolation + generalised eigenproblem + occupations + rdm1 + one full Fock build)."""
    import graddft_b200 as gd
    from graddft_b200 import ops

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    dgemm_tf, _ = ctx.dgemm()
    hbm_gbs = hbm_peak()
    sh = SCF_SHAPES[shape_key]
    N, n = sh["N"], sh["n"]
    m = _scf_shard(N, n, rank, world, dev, n_omega=1)
    lo_c, hi_c = 2, 2 + steps
    loops = {c: gd.make_jitted_scf_loop(gd.B3LYP, cycles=c) for c in (lo_c, hi_c)}
    eager = {c: gd.diff_scf_loop(gd.B3LYP, cycles=c) for c in (lo_c, hi_c)}
    with torch.no_grad():
        for c in (lo_c, hi_c):
            out = loops[c](None, m)  # warm-up + capture (workspaces, handles)
            eager[c](None, m)
        ms, ms_eager = {}, {}
        for c in (lo_c, hi_c):
            ms[c] = min(ctx.timed(lambda: loops[c](None, m), 1) for _ in range(3))
            ms_eager[c] = min(ctx.timed(lambda: eager[c](None, m), 1) for _ in range(2))
        per_iter = (ms[hi_c] - ms[lo_c]) / steps
        launches0 = ops.lib().gdft_launch_count()
        ops.TIMING = {}
        out = eager[lo_c](None, m)
        torch.cuda.synchronize()
        timing, ops.TIMING = ops.TIMING, None
        launches = (ops.lib().gdft_launch_count() - launches0) / (lo_c + 1.0)
        e_final, f_final = out.energy.clone(), out.fock.clone()
        e2e = None
        if with_e2e:
            # the whole public call with the initial density matrix arriving from pinned host memory and (energy, Fock
            # matrix, rdm1) returned to pinned host memory: iterations/s of a complete (2 + steps)-cycle call, the initial
            # Fock build and the copies included
            rd_host = m.rdm1.cpu().pin_memory()
            res_host = torch.empty(1 + 4 * n * n, dtype=torch.float64).pin_memory()
            res_dev = torch.empty(1 + 4 * n * n, dtype=torch.float64, device=dev)

            def call():
                m.rdm1.copy_(rd_host, non_blocking=True)
                o = loops[hi_c](None, m)
                res_dev[0] = o.energy
                res_dev[1:1 + 2 * n * n] = o.fock.reshape(-1)
                res_dev[1 + 2 * n * n:] = o.rdm1.reshape(-1)
                res_host.copy_(res_dev, non_blocking=True)

            call()
            ms_call = min(ctx.timed(call, 1) for _ in range(3))
            e2e = {"value": hi_c / (ms_call / 1e3), "unit": "iter/s", "h2d_bytes_per_step": rd_host.numel() * 8 / hi_c,
                   "d2h_bytes_per_step": res_host.numel() * 8 / hi_c, "ms_per_call": ms_call, "cycles_per_call": hi_c,
                   "note": "iterations/s of one complete make_jitted_scf_loop call (initial Fock build + cycles) with rdm1 from pinned host "
                           "memory and (E, Fock, rdm1) back to pinned host memory inside the timed region; bytes are per iteration"}

    def avg(name):
        ev = timing.get(name, [])
        return sum(a.elapsed_time(b) for a, b in ev) / max(1, len(ev))

    graphed = bool(getattr(loops[hi_c], "last_call_was_graph", False))
    res = {"workload": sh["desc"], "N": N, "n": n, "iter_per_s": 1e3 / per_iter, "ms_per_iter": per_iter,
           f"ms_loop_{lo_c}_cycles": ms[lo_c], f"ms_loop_{hi_c}_cycles": ms[hi_c], "eager_ms_per_iter": (ms_eager[hi_c] - ms_eager[lo_c]) / steps,
           "cuda_graph": graphed, "energy_finite": bool(torch.isfinite(e_final)), "gpu_launches_per_fock_build": launches}
    if e2e is not None:
        res["e2e"] = e2e
    eri_numel = m.rep_tensor.numel()
    # bytes one J sweep reads: the packed pair block when the tensor was packed (a quarter of the full rows), else the block as given
    from graddft_b200 import distributed as _gd
    _pe = (ops.packed_eri_for(m.rep_tensor, count_use=False) if world == 1 else
           next((e["packed"] for e in _gd._PACKED_BLOCKS if e["ref"]() is m.rep_tensor), None))
    eri_sweep_bytes = 8.0 * (_pe.pairs * (_pe.packed.numel() // max(_pe.pairs, 1)) if _pe else eri_numel)
    eri_packed = bool(_pe)
    del _pe
    ops.release_packed_eri()
    del m, out
    torch.cuda.empty_cache()
    if with_parity and world > 1:
        parity = None
        if rank == 0:
            mu = _scf_unsharded(N, n, world, dev, 1)
            with torch.no_grad():
                o = gd.diff_scf_loop(gd.B3LYP, cycles=lo_c)(None, mu)
            dE = abs(float(o.energy - e_final))
            dF = float((o.fock - f_final).abs().max() / o.fock.abs().max())
            parity = {"against": f"rank-0 single-GPU {lo_c}-cycle diff_scf_loop on the unsharded molecule", "dE_vs_1gpu": dE, "dFock_rel": dF,
                      "E": float(e_final), "tolerance": {"dE": E_TOL, "dFock_rel": V_RTOL}, "ok": bool(dE < E_TOL and dF < V_RTOL)}
            del mu, o
            torch.cuda.empty_cache()
        ctx.barrier()
        if parity is not None:
            res["parity"] = parity
            if not parity["ok"]:
                ctx.fail(f"scf_{shape_key} x{world}", parity)
    if rank == 0:
        eri_ms = avg("gdft_eri_jk")
        eri_bytes = eri_sweep_bytes
        res["kernels_ms"] = {"density_fwd": avg("gdft_density_fwd"), "density_bwd": avg("gdft_density_bwd"), "eri_j": eri_ms,
                             "sym_eigh": avg("gdft_sym_eigh")}
        if eri_bytes > 2.5e8:  # larger than L2: a DRAM figure
            res["roofline_eri"] = {"bound": "hbm", "kernel": "eri_j_kernel (rep_tensor (pq)x(rt) sweep)", "achieved": eri_bytes / eri_ms / 1e6,
                                   "peak": hbm_gbs[0], "unit": "GB/s", "frac": eri_bytes / eri_ms / 1e6 / hbm_gbs[0],
                                   "bytes_per_launch": eri_bytes, "ms_per_launch": eri_ms, "peak_source": hbm_gbs[1], "packed": eri_packed,
                                   "unpacked_bytes": 8.0 * n ** 4 / world}
        # roofline of one iteration: B3LYP Fock build = 16 GEMM units (rho, grad, lapl fwd + VJP) + 2 (HF Fock) at the
        # measured DGEMM rate, plus one rep_tensor sweep at the HBM peak (per-GPU shares)
        unit = 2.0 * (N / world) * n * n
        t_roof = 18.0 * unit / (dgemm_tf * 1e9) + eri_bytes / (hbm_gbs[0] * 1e6)
        res["roofline_iter"] = {"ms_at_roofline": t_roof, "frac": t_roof / per_iter,
                                "model": "18 units x 2*N*n^2 FLOP at measured cuBLAS DGEMM + the bytes one J sweep reads (packed pair rows: 8*pairs*npair B; "
                                         "un-packed: 8*rows*n^2 B) at HBM peak"}
    return res


def dm21_leg(ctx, steps=3, with_parity=True, with_e2e=False):
    """BASELINE configs[2]: DM21 neural functional energy + gradient (one energy_predictor call = E and the Fock matrix)."""
    import graddft_b200 as gd

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    sh = SCF_SHAPES["c3_dm21"]
    N, n = sh["N"], sh["n"]
    m = _scf_shard(N, n, rank, world, dev, n_omega=2)
    functional = gd.DM21()
    params = functional.generate_DM21_weights(device=dev)
    pred = gd.energy_predictor(functional)
    with torch.no_grad():
        for _ in range(2):
            e, f = pred(params, m)
        ms_pred = min(ctx.timed(lambda: pred(params, m), steps) / steps for _ in range(2))
        e2e = None
        if with_e2e:
            rd_host = m.rdm1.cpu().pin_memory()
            res_host = torch.empty(1 + 2 * n * n, dtype=torch.float64).pin_memory()
            res_dev = torch.empty(1 + 2 * n * n, dtype=torch.float64, device=dev)

            def call():
                m.rdm1.copy_(rd_host, non_blocking=True)
                ee, ff = pred(params, m)
                res_dev[0] = ee
                res_dev[1:] = ff.reshape(-1)
                res_host.copy_(res_dev, non_blocking=True)

            call()
            ms_call = min(ctx.timed(call, steps) / steps for _ in range(2))
            e2e = {"value": 1e3 / ms_call, "unit": "calls/s", "h2d_bytes_per_step": rd_host.numel() * 8, "d2h_bytes_per_step": res_host.numel() * 8,
                   "ms_per_step": ms_call}
    res = {"workload": sh["desc"], "N": N, "n": n, "predict_ms": ms_pred, "predicts_per_s": 1e3 / ms_pred,
           "energy_finite": bool(torch.isfinite(e)) and bool(torch.isfinite(f).all())}
    if e2e is not None:
        res["e2e"] = e2e
    e_s, f_s = e.clone(), f.clone()
    del m, e, f
    torch.cuda.empty_cache()
    if with_parity and world > 1:
        parity = None
        if rank == 0:
            mu = _scf_unsharded(N, n, world, dev, 2)
            with torch.no_grad():
                e1, f1 = pred(params, mu)
            dE = abs(float(e1 - e_s))
            dF = float((f1 - f_s).abs().max() / f1.abs().max())
            parity = {"against": "rank-0 single-GPU energy_predictor call on the unsharded molecule", "dE_vs_1gpu": dE, "dFock_rel": dF, "E": float(e_s),
                      "tolerance": {"dE": E_TOL, "dFock_rel": V_RTOL}, "ok": bool(dE < E_TOL and dF < V_RTOL)}
            del mu, e1, f1
            torch.cuda.empty_cache()
        ctx.barrier()
        if parity is not None:
            res["parity"] = parity
            if not parity["ok"]:
                ctx.fail(f"dm21_c3 x{world}", parity)
    return res


def _train_shapes():
    g = torch.Generator().manual_seed(1993)
    return [(int(1e4 * (1 + 3 * float(torch.rand((), generator=g)))), 12 + int(88 * float(torch.rand((), generator=g)))) for _ in range(64)]


def train_leg(ctx, with_parity=True, with_e2e=False):
    """BASELINE configs[4]: one training step (non-SCF energy loss, grad_dft/train.py:312-359,480-535) of the DM21-shaped
    neural functional on a batch of 64 small synthetic molecules (n_i = 12 + floor(88 U), N_i = 1e4 (1 + 3 U), seed
    1993), molecules sharded over the ranks (balanced by N n^2), one gradient all-reduce per step, Adam update."""
    import graddft_b200 as gd
    from graddft_b200 import distributed as gdist
    from graddft_b200.synthetic import synthetic_molecule

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    shapes = _train_shapes()

    def make(i):
        m = gd.molecule_from_tensors(synthetic_molecule(shapes[i][0], shapes[i][1], n_omega=2, seed=1993 + i, device=dev, mask_frac=0.0), dev)
        m.packed_basis
        return m

    mine = gdist.shard_molecules([N * n * n for N, n in shapes], rank, world)
    mols = {i: make(i) for i in mine}
    fun = gd.DM21()
    params = {k: v.requires_grad_(True) for k, v in fun.generate_DM21_weights(device=dev).items()}
    leaves = list(params.values())
    opt = torch.optim.Adam(leaves, lr=1e-4)
    predictor = gd.non_scf_predictor(fun)
    truth = lambda i: torch.tensor(-1.0 - 0.01 * i, dtype=torch.float64, device=dev)  # noqa: E731
    truths = {i: truth(i) for i in range(64)}

    def loss_and_grads(idx, molecules):
        total = torch.zeros((), dtype=torch.float64, device=dev)
        # what mse_energy_loss evaluates (the loss reads .energy only): features per molecule, one network pass per group
        energies = predictor.energy_only_batch(params, [molecules[i] for i in idx])
        for i, energy in zip(idx, energies):
            total = total + ((energy - truths[i]) / molecules[i].mo_occ.sum()) ** 2
        loss = total / 64
        return loss.detach(), torch.autograd.grad(loss, leaves)

    def step():
        loss, grads = loss_and_grads(mine, mols)
        grads, loss = gdist.allreduce_gradients(list(grads), loss)
        for p_, g_ in zip(leaves, grads):
            p_.grad = g_
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    parity = None
    if with_parity and world > 1:
        loss_s, grads_s = loss_and_grads(mine, mols)
        grads_s, loss_s = gdist.allreduce_gradients(list(grads_s), loss_s)
        if rank == 0:
            loss_1 = torch.zeros((), dtype=torch.float64, device=dev)
            grads_1 = [torch.zeros_like(p_) for p_ in leaves]
            for i0 in range(0, 64, 8):  # all 64 molecules on this one GPU, eight at a time
                idx = list(range(i0, i0 + 8))
                ms_ = {i: (mols[i] if i in mols else make(i)) for i in idx}
                l, gs = loss_and_grads(idx, ms_)
                loss_1 += l
                for a, b in zip(grads_1, gs):
                    a += b
                del ms_
            dL = abs(float(loss_1 - loss_s)) / max(1e-300, abs(float(loss_1)))
            scale = max(float(g_.abs().max()) for g_ in grads_1)
            dG = max(float((a - b).abs().max()) for a, b in zip(grads_s, grads_1)) / scale
            parity = {"against": "rank-0 single-GPU loss and parameter gradient over all 64 molecules", "dLoss_rel": dL, "dGrad_rel": dG,
                      "loss": float(loss_s), "tolerance": {"dLoss_rel": V_RTOL, "dGrad_rel": V_RTOL}, "ok": bool(dL < V_RTOL and dG < V_RTOL)}
        ctx.barrier()
        if parity is not None and not parity["ok"]:
            ctx.fail(f"train_c5 x{world}", parity)
    for _ in range(2):
        loss = step()
    ms = min(ctx.timed(step, 1) for _ in range(3))
    res = {"workload": TRAIN_DESC, "ms_per_step": ms, "molecules_per_s": 64e3 / ms, "molecules_on_rank0": len(mine), "loss_finite": bool(torch.isfinite(loss))}
    if with_e2e:
        # per step the rank's density matrices arrive from pinned host memory and the loss goes back
        hosts = {i: mols[i].rdm1.cpu().pin_memory() for i in mine}
        loss_host = torch.empty(1, dtype=torch.float64).pin_memory()

        def step_e2e():
            for i in mine:
                mols[i].rdm1.copy_(hosts[i], non_blocking=True)
            loss_host.copy_(step().reshape(1), non_blocking=True)

        step_e2e()
        ms_e = min(ctx.timed(step_e2e, 1) for _ in range(3))
        res["e2e"] = {"value": 1e3 / ms_e, "unit": "steps/s", "h2d_bytes_per_step": sum(h.numel() for h in hosts.values()) * 8, "d2h_bytes_per_step": 8,
                      "ms_per_step": ms_e}
    if parity is not None:
        res["parity"] = parity
    return res


def chi_leg(ctx):
    """Row f4: the chi-generation tail at the benzene shape (n = 264): `rows` grid points per GPU with their nu[r] (n x n
    per point, 557 KB) resident in HBM, streamed once by gdft_chi_contract; roofline = 8 n^2 bytes per point against the
    HBM peak.  `e2e` (1 GPU only): the same points with nu arriving from host memory in 1024-point chunks through the
    double-buffered uploader of generate_chi_tensor (PCIe-bound by construction: nu is produced on the host by libcint)."""
    from graddft_b200 import interface, ops

    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    n, rows, chunk = 264, 9472, 1024  # 148 SMs x 2 CTAs x 8 points x 4 groups
    g = torch.Generator(device=dev).manual_seed(1984 + rank)
    ao = torch.randn(rows, n, generator=g, dtype=torch.float64, device=dev)
    D = torch.randn(2, n, n, generator=g, dtype=torch.float64, device=dev)
    nu = torch.randn(rows, n, n, generator=g, dtype=torch.float64, device=dev)  # 5.3 GB >> L2
    chi = torch.empty((rows, 1, 2, n), dtype=torch.float64, device=dev)

    def step():
        ops.chi_contract_(chi, 0, 0, ao, D, nu)

    for _ in range(3):
        step()
    steps = 10
    ms = ctx.timed(step, steps) / steps
    bytes_per_launch = 8.0 * rows * n * n
    hbm_gbs, hbm_src = hbm_peak()
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
        ms_e = ctx.timed(e2e, 3) / 3
        out["e2e"] = {"value": host_rows / (ms_e / 1e3), "unit": "points/s", "h2d_bytes_per_step": 8 * host_rows * n * n,
                      "d2h_bytes_per_step": 0, "ms_per_step": ms_e, "h2d_GBps": 8e-6 * host_rows * n * n / ms_e,
                      "note": "nu chunks of 1024 points from pageable host memory, staged into two cached page-locked buffers by a pool of host threads, "
                              "each row slice uploaded as soon as it is staged"}
        nu_pinned = nu_host.pin_memory()

        def e2e_pinned():
            interface.generate_chi_tensor(D, ao[:host_rows], cidx, lambda c, omega: nu_pinned[int(c[0, 0]):int(c[0, 0]) + len(c)], [0.0], chunk)

        e2e_pinned()
        ms_p = ctx.timed(e2e_pinned, 3) / 3
        out["e2e_pinned_source"] = {"value": host_rows / (ms_p / 1e3), "unit": "points/s", "ms_per_step": ms_p,
                                    "h2d_GBps": 8e-6 * host_rows * n * n / ms_p,
                                    "note": "the provider hands over page-locked chunks: no staging copy, H2D overlapped with the kernel"}
    return out


def extras(ctx, args):
    """The secondary legs of the default run (kept under the key "scf" of the c4 line)."""
    scf = {}

    def guarded(name, fn):
        try:
            scf[name] = fn()
        except SystemExit:
            raise
        except Exception as exc:  # the headline XC line must survive a failure of a secondary leg
            scf[name] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    if ctx.world == 1:
        guarded("c2", lambda: scf_leg(ctx, "c2"))
    guarded("c3", lambda: scf_leg(ctx, "c3"))
    guarded("c3_dm21", lambda: dm21_leg(ctx))
    guarded("c5_training", lambda: train_leg(ctx))
    guarded("chi_tail", lambda: chi_leg(ctx))
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu_baseline:
        if "error" not in scf.get("c2", {"error": 1}):
            sh = SCF_SHAPES["c2"]
            r = cpu_scf_iteration(sh["N"], sh["n"], sh["N"], sh["n"], reps=2)
            scf["c2"]["cpu_baseline"] = {"value": 1.0 / r["s_per_iter"], "unit": "iter/s", "cores": r["cores"], "kind": "port",
                                         "sample": f"full H2O-shaped molecule, oracle diff_scf_loop iteration (torch-CPU float64), {r['s_per_iter'] * 1e3:.0f} ms/iter"}
        if "error" not in scf.get("c3", {"error": 1}):
            sh = SCF_SHAPES["c3"]
            r = cpu_scf_iteration(sh["N"], sh["n"], 25_000, 16)
            scf["c3"]["cpu_baseline"] = {"value": 1.0 / r["s_per_iter"], "unit": "iter/s", "cores": r["cores"], "kind": "port",
                                         "sample": f"oracle diff_scf_loop iteration composed of the grid part on {r['grid_rows']} of {sh['N']} rows "
                                                   f"({r['grid_s']:.2f} s), a rep_tensor sweep on {r['eri_p']} of {sh['n']} slabs ({r['eri_s']:.3f} s) and the "
                                                   f"n x n part in full ({r['small_s']:.3f} s), scaled linearly: {r['s_per_iter']:.2f} s/iter"}
    return scf


def run_secondary(ctx, args, key):
    """--workload scf_c2 | scf_c3 | dm21_c3 | train_c5: the leg as the line's own metric."""
    rank, world = ctx.rank, ctx.world
    sampler = ClockSampler(ctx.local_rank)
    if rank == 0:
        sampler.start()
    steps = max(1, args.steps)
    if key in ("scf_c2", "scf_c3"):
        sk = key[4:]
        res = scf_leg(ctx, sk, steps=steps, with_e2e=True)
        line = {"metric": "jitted_scf_iter_per_s", "value": res["iter_per_s"], "unit": "iter/s", "ms_per_step": res["ms_per_iter"], "config": _scf_config(sk, world)}
        if rank == 0:
            dominant = max(res["kernels_ms"], key=lambda k: res["kernels_ms"][k])
            if "roofline_eri" in res and sk == "c3":
                line["roofline"] = dict(res["roofline_eri"], traffic=None)
            line["roofline_iter"] = res["roofline_iter"]
            line["dominant_kernel"] = dominant
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            sh = SCF_SHAPES[sk]
            full = sk == "c2"
            r = cpu_scf_iteration(sh["N"], sh["n"], sh["N"] if full else 25_000, sh["n"] if full else 16, reps=2)
            cpu = {"value": 1.0 / r["s_per_iter"], "unit": "iter/s", "cores": r["cores"], "kind": "port",
                   "sample": f"oracle diff_scf_loop iteration: grid part on {r['grid_rows']} of {sh['N']} rows ({r['grid_s']:.3f} s), rep_tensor sweep on "
                             f"{r['eri_p']} of {sh['n']} slabs ({r['eri_s']:.3f} s), n x n part in full ({r['small_s']:.3f} s), scaled linearly: "
                             f"{r['s_per_iter']:.3f} s/iter"}
    elif key == "dm21_c3":
        res = dm21_leg(ctx, steps=steps, with_e2e=True)
        line = {"metric": "dm21_energy_predictor_calls_per_s", "value": res["predicts_per_s"], "unit": "calls/s", "ms_per_step": res["predict_ms"],
                "config": _scf_config("c3_dm21", world)}
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            sh = SCF_SHAPES["c3_dm21"]
            r = cpu_dm21_predict(sh["N"], sh["n"], 10_000, 16)
            cpu = {"value": 1.0 / r["s_per_call"], "unit": "calls/s", "cores": r["cores"], "kind": "port",
                   "sample": f"oracle predict_dm21 pieces: grid part on {r['grid_rows']} of {sh['N']} rows ({r['grid_s']:.2f} s), rep_tensor sweep on "
                             f"{r['eri_p']} of {sh['n']} slabs ({r['eri_s']:.3f} s), scaled linearly: {r['s_per_call']:.1f} s/call"}
    else:
        res = train_leg(ctx, with_e2e=True)
        line = {"metric": "training_steps_per_s", "value": 1e3 / res["ms_per_step"], "unit": "steps/s", "ms_per_step": res["ms_per_step"], "config": _train_config(world)}
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            r = cpu_train_step(_train_shapes(), [0, 21, 42, 63])
            cpu = {"value": 1.0 / r["s_per_step"], "unit": "steps/s", "cores": r["cores"], "kind": "port",
                   "sample": f"oracle energy + parameter gradient of molecules {r['sample']} of 64 ({r['sample_s']:.2f} s), scaled by N (n^2 + 1e5): {r['s_per_step']:.1f} s/step"}
    clocks = sampler.stop() if rank == 0 else None
    if rank != 0:
        return
    line.update({"n_gpus": world, "steps": steps, "warmup": 2, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                 "data": "synthetic", "clocks": clocks, "e2e": res.pop("e2e", None), "detail": res})
    if "parity" in res:
        line["parity"] = res["parity"]
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(json.dumps(line))


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
    ap.add_argument("--workload", default="c4", choices=WORKLOADS)
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--no-extras", "--no-scf", dest="no_extras", action="store_true", help="skip the secondary legs of the default run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    ctx = Ctx()
    if args.workload in XC_SHAPES:
        run_xc(ctx, args, args.workload, None if (args.no_extras or args.workload != "c4") else extras)
    else:
        run_secondary(ctx, args, args.workload)
    ctx.close()


if __name__ == "__main__":
    main()
