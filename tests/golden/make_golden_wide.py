"""tests/golden/predictor_wide.npz: energy_predictor of the reference's own source files (imported from /root/reference,
unmodified, on the torch-backed jax stand-in of jaxshim.py, like make_golden.py) at the widths BASELINE.json names for the
small configurations -- 43 AOs (H2O / def2-TZVP) and 97 AOs -- where the CUDA kernels run other tile classes than at the n <= 12
of predictor_{a,b}.npz -- and tests/golden/scf_wide.npz: diff_scf_loop (B3LYP, 3 DIIS cycles; B88, 5) and the DM21 predictor at 43 AOs; and tests/golden/train_batch.npz: mse_energy_loss of a three-molecule batch with its parameter gradient; tests/golden/scf_grad.npz: the parameter gradient THROUGH two SCF cycles.  The inputs are NOT stored (a 97^4 rep_tensor is 708 MB): they are `synthetic_molecule(N, n, seed)`,
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


def gapped(mol):
    """A well-separated orbital spectrum, so that aufbau occupations do not switch inside the loop (the SCF map stays smooth in
    the parameters); tests/test_predictor_gpu.py::_gapped_molecule is the same construction."""
    n = mol["h1e"].shape[-1]
    mol["h1e"] = torch.diag(torch.linspace(-8.0, 8.0, n, dtype=torch.float64)) + 0.05 * mol["h1e"]
    mol["rep_tensor"] = 0.05 * mol["rep_tensor"]
    mol["s1e"] = torch.eye(n, dtype=torch.float64) + 0.2 * (mol["s1e"] - torch.eye(n, dtype=torch.float64))
    return mol


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

    # ---- the DIIS loop (evaluate.py:917-1038) and the DM21 predictor at 43 AOs -> scf_wide.npz ------------------------------
    import oracle

    d = {}
    N, n, seed = 3000, 43, 2143
    mol = mg.synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0)
    m = mg.ref_molecule(mol)
    d["shape"], d["checksums"] = np.array([N, n, seed]), checksums(mol)
    for name, cycles in (("B3LYP", 3), ("B88", 5)):
        out = mg.gd.diff_scf_loop(getattr(mg.gd, name), cycles=cycles)(None, m)
        d[f"diis_energy_{name}_{cycles}"], d[f"diis_rdm1_{name}_{cycles}"], d[f"diis_fock_{name}_{cycles}"] = mg.np_(out.energy), mg.np_(out.rdm1), mg.np_(out.fock)
        print("scf", name, cycles, float(out.energy))
    flat = oracle.dm21_mlp_init(width=32, n_layers=3, seed=2143)
    tree = {}
    for k, v in flat.items():
        layer, leaf = k.split(".")
        tree.setdefault(layer, {})[leaf] = mg.J(v)
    dm21 = mg.gd.DM21()
    dm21.layer_widths = [32, 32, 32]
    e, fock = mg.gd.energy_predictor(dm21)({"params": tree}, m)
    for k, v in flat.items():
        d["param_" + k] = mg.np_(v)
    d["energy_DM21"], d["fock_DM21"] = mg.np_(e), mg.np_(fock)
    print("DM21", float(e))
    np.savez_compressed(HERE / "scf_wide.npz", **d)
    print((HERE / "scf_wide.npz").stat().st_size)

    # ---- training batch (train.py:480-535 mse_energy_loss over non_scf_predictor, evaluate.py:88-126) -> train_batch.npz ------
    # the loss of the reference's own source and its gradient w.r.t. the network parameters (torch autograd through the stand-in,
    # where jax.grad would act), for three molecules of different sizes and a DM21-shaped network with seeded weights
    tr = sys.modules["grad_dft.train"]
    d = {}
    shapes = [(900, 9, 3101), (1200, 14, 3102), (700, 7, 3103)]
    atom_index = [[8, 1, 1], [6, 1, 1, 1, 1], [7, 1, 1, 1]]
    mols = [mg.synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0) for N, n, seed in shapes]
    ms = [mg.ref_molecule(mol).replace(atom_index=mg.J(torch.tensor(z, dtype=torch.int64))) for mol, z in zip(mols, atom_index)]
    d["shapes"] = np.array(shapes)
    d["atom_index"] = np.array([z + [0] * (5 - len(z)) for z in atom_index])
    d["checksums"] = np.stack([checksums(mol) for mol in mols])
    truths = torch.tensor([-70.0, -40.0, -55.0], dtype=torch.float64)
    d["truths"] = mg.np_(truths)
    flat = oracle.dm21_mlp_init(width=32, n_layers=3, seed=3100)
    leaves = {k: mg.J(v.clone().requires_grad_(True)) for k, v in flat.items()}
    tree = {}
    for k, v in leaves.items():
        layer, leaf = k.split(".")
        tree.setdefault(layer, {})[leaf] = v
    dm21 = mg.gd.DM21()
    dm21.layer_widths = [32, 32, 32]
    predictor = sys.modules["grad_dft.evaluate"].non_scf_predictor(dm21)
    with torch.enable_grad():
        for norm in (True, False):
            loss = tr.mse_energy_loss({"params": tree}, predictor, ms, mg.J(truths), norm)
            grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
            tag = "norm" if norm else "plain"
            d[f"loss_{tag}"] = mg.np_(loss)
            for k, g in zip(leaves, grads):
                d[f"grad_{tag}_{k}"] = mg.np_(g if g is not None else torch.zeros_like(leaves[k]))
            print("train", tag, float(loss))
    for k, v in flat.items():
        d["param_" + k] = mg.np_(v)
    np.savez_compressed(HERE / "train_batch.npz", **d)
    print((HERE / "train_batch.npz").stat().st_size)

    # ---- training THROUGH the SCF loop (evaluate.py:917-1038 under jax.grad; examples/advanced_scripts/train_scf_loop.py)
    # -> scf_grad.npz: energy after 2 DIIS cycles / 2 linear-mixing cycles of the hybrid DM21 functional and its gradient
    # w.r.t. the parameters, from the reference's own loops (stop_gradients, safe-eigh custom VJP, DIIS included).  The molecule
    # is synthetic_molecule with a well-separated orbital spectrum (`gapped`, below; the tests apply the same three lines).
    d = {}
    N, n, seed = 700, 6, 3301
    mol = gapped(mg.synthetic_molecule(N, n, n_omega=2, seed=seed, mask_frac=0.0))
    d["shape"], d["checksums"] = np.array([N, n, seed]), checksums(mol)
    m = mg.ref_molecule(mol)
    flat = oracle.dm21_mlp_init(width=8, n_layers=2, seed=3300)
    dm21 = mg.gd.DM21()
    dm21.layer_widths = [8, 8]
    ev = sys.modules["grad_dft.evaluate"]
    for tag, make in (("diis", lambda: ev.diff_scf_loop(dm21, cycles=2)), ("simple", lambda: ev.diff_simple_scf_loop(dm21, cycles=2))):
        leaves = {k: mg.J(v.clone().requires_grad_(True)) for k, v in flat.items()}
        tree = {}
        for k, v in leaves.items():
            layer, leaf = k.split(".")
            tree.setdefault(layer, {})[leaf] = v
        with torch.enable_grad():
            out = make()({"params": tree}, m)
            grads = torch.autograd.grad(out.energy, list(leaves.values()), allow_unused=True)
        d[f"energy_{tag}"] = mg.np_(out.energy)
        for k, g in zip(leaves, grads):
            d[f"grad_{tag}_{k}"] = mg.np_(g if g is not None else torch.zeros_like(leaves[k]))
        print("scf grad", tag, float(out.energy), max(float(g.abs().max()) for g in grads if g is not None))
    for k, v in flat.items():
        d["param_" + k] = mg.np_(v)
    np.savez_compressed(HERE / "scf_grad.npz", **d)
    print((HERE / "scf_grad.npz").stat().st_size)


if __name__ == "__main__":
    main()
