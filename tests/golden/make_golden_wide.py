"""tests/golden/predictor_wide.npz: energy_predictor of the reference's own source files (imported from /root/reference,
unmodified, on the torch-backed jax stand-in of jaxshim.py, like make_golden.py) at the widths BASELINE.json names for the
small configurations -- 43 AOs (H2O / def2-TZVP) and 97 AOs -- where the CUDA kernels run other tile classes than at the n <= 12
of predictor_{a,b}.npz.  The inputs are NOT stored (a 97^4 rep_tensor is 708 MB): they are `synthetic_molecule(N, n, seed)`,
which the test regenerates; a few input checksums are stored so that a drifting generator fails loudly instead of silently
comparing different molecules.

    python tests/golden/make_golden_wide.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import make_golden as mg  # noqa: E402  (installs the stand-in and imports the reference package)

CASES = {"n43": dict(N=3000, n=43, seed=2043, names=("LSDA", "B88", "VWN", "LYP", "PW92", "B3LYP")),
         "n97": dict(N=2000, n=97, seed=2097, names=("B88", "B3LYP"))}


def checksums(mol):
    return np.array([float(mol[k].double().sum()) for k in ("ao", "grad_ao", "rdm1", "weights", "rep_tensor", "chi", "h1e")])


def main():
    d = {}
    for tag, c in CASES.items():
        mol = mg.synthetic_molecule(c["N"], c["n"], n_omega=2, seed=c["seed"], mask_frac=0.0)
        m = mg.ref_molecule(mol)
        d[f"{tag}_shape"] = np.array([c["N"], c["n"], c["seed"]])
        d[f"{tag}_checksums"] = checksums(mol)
        for name in c["names"]:
            e, fock = mg.gd.energy_predictor(getattr(mg.gd, name))(None, m)
            d[f"{tag}_energy_{name}"], d[f"{tag}_fock_{name}"] = mg.np_(e), mg.np_(fock)
            print(tag, name, float(e))
    np.savez_compressed(HERE / "predictor_wide.npz", **d)
    print((HERE / "predictor_wide.npz").stat().st_size)


if __name__ == "__main__":
    main()
