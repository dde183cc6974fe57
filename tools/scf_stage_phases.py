"""Per-phase clocks of scf_diis_kernel from a -DGDFT_SCF_PROF build of the library (development only; the build is made
here on first use into tools/_dev/, which is git-ignored)."""
import ctypes, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from graddft_b200 import _lib
from graddft_b200.build import build as _build, NVCC_FLAGS, OBJ, CSRC
DEV = Path(__file__).resolve().parent / "_dev"
if not (DEV / "libgdft_prof.so").exists():
    _build()
    DEV.mkdir(exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")]
    subprocess.run(["nvcc", *flags, "-DGDFT_SCF_PROF", "-c", str(CSRC / "scf_glue.cu"), "-o", str(DEV / "scf_glue_prof.o")], check=True)
    objs = [str(o) for o in OBJ.glob("*.o") if o.name != "scf_glue.o"]
    subprocess.run(["nvcc", "-shared", "-o", str(DEV / "libgdft_prof.so"), str(DEV / "scf_glue_prof.o"), *objs,
                    "-gencode", "arch=compute_100a,code=sm_100a", "-ldl"], check=True)
_lib.LIB_PATH = DEV / "libgdft_prof.so"
from graddft_b200 import ops, evaluate
dev = torch.device("cuda:0"); F64 = torch.float64
names = ["load", "FD,(FD)S", "err/fock store", "gram", "B build", "LU+solve", "combine+L load", "2 matmuls + store"]
for n in (43, 64):
    m = 10
    g = torch.Generator(device=dev).manual_seed(1)
    rn = lambda *s: torch.randn(*s, generator=g, dtype=F64, device=dev)
    X = rn(n, n); S = X @ X.T / n + torch.eye(n, dtype=F64, device=dev); L_inv = evaluate.overlap_factor(S)
    z = torch.zeros((m, 2, n, n), dtype=F64, device=dev); fv, ev, gram = z.clone(), z.clone(), torch.zeros((2, m, m), dtype=F64, device=dev)
    F, D = rn(2, n, n), rn(2, n, n)
    for c in range(14): ops.scf_diis_step(c, F + 0.1 * rn(2, n, n), D, S, L_inv, fv, ev, gram)
    _, _, x = ops.scf_diis_step(14 | (1 << 16), F, D, S, L_inv, fv, ev, gram)
    torch.cuda.synchronize()
    st = x.reshape(-1)[:9].tolist()
    print(f"n={n}: total {st[-1]:.0f} clocks = {st[-1] / 1.965e3:.1f} us")
    for i, nm in enumerate(names): print(f"   {nm:22s} {st[i + 1] - st[i]:8.0f} clocks")
