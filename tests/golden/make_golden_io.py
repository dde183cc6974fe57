"""Generates tests/golden/io_chi.npz, io_tree.npz and io_loaded.npz by executing the
reference's own, unmodified grad_dft/interface/pyscf.py (loaded from /root/reference) on the torch-backed jax stand-in
(jaxshim.py), permissive stubs for the PySCF modules it imports at module level (none of their code runs), and the
in-memory h5py stand-in (fake_h5py.py):

  io_chi.npz     inputs (ao, rdm1, seeded nu chunks) and the output of generate_chi_tensor (pyscf.py:1067-1124) with
                 `_nu_chunk` replaced by the seeded provider (nu comes from libcint upstream; out of path);
  io_tree.npz    the flattened group/dataset/attribute tree that saver (pyscf.py:331-426) writes for two molecules and
                 one reaction;
  io_loaded.npz  what loader (pyscf.py:430-581) yields from that tree for the argument combinations the tests use.

    python tests/golden/make_golden_io.py
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import fake_h5py  # noqa: E402
import jaxshim  # noqa: E402
from graddft_b200.synthetic import synthetic_molecule  # noqa: E402

F64 = torch.float64
gd = jaxshim.install()
J = jaxshim._j
REF = Path("/root/reference")


class _Anything(types.ModuleType):
    """A module whose every attribute is a dummy class (only ever used in annotations and isinstance checks)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = sys.modules.get(f"{self.__name__}.{name}")
        if sub is not None:
            return sub
        return type(name, (), {})


def stub(name):
    m = _Anything(name)
    m.__path__ = []
    sys.modules[name] = m
    return m


for nm in ("pyscf", "pyscf.scf", "pyscf.dft", "pyscf.pbc", "pyscf.pbc.dft", "pyscf.gto", "pyscf.data", "pyscf.data.elements",
           "pyscf.pbc.gto", "pyscf.pbc.gto.cell", "pyscf.pbc.lib", "pyscf.pbc.lib.kpts", "pyscf.pbc.df", "pyscf.pbc.df.fft",
           "pyscf.pbc.df.mdf", "pyscf.pbc.df.df", "pyscf.ao2mo", "pyscf.cc", "grad_dft.solid", "grad_dft.external"):
    stub(nm)
sys.modules["grad_dft.external"]._nu_chunk = None
sys.modules["h5py"] = fake_h5py
utils = sys.modules["grad_dft.utils"]
for nm in ("DType", "DensityFunctional", "HartreeFock"):
    setattr(utils, nm, type(nm, (), {}))
utils.default_dtype = lambda: F64
sys.modules["jaxtyping"].Bool = sys.modules["jaxtyping"].Int


def real_tree_map(f, tree):
    """jax.tree_util.tree_map on dicts: recursive, None is an empty subtree."""
    if tree is None:
        return None
    if isinstance(tree, dict):
        return {k: real_tree_map(f, v) for k, v in tree.items()}
    return f(tree)


sys.modules["jax.tree_util"].tree_map = real_tree_map
molmod = sys.modules["grad_dft.molecule"]
spec = importlib.util.spec_from_file_location("grad_dft.interface.pyscf_reference", REF / "grad_dft" / "interface" / "pyscf.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


class _JnpForDatasets:
    """jax.numpy as the loader sees it: jnp.asarray / jnp.array accept an h5py dataset through __array__."""

    def __init__(self, base):
        self._base = base

    def __getattr__(self, name):
        return getattr(self._base, name)

    def asarray(self, x, dtype=None):
        return self._base.asarray(np.asarray(x) if isinstance(x, fake_h5py.Dataset) else x, dtype=dtype)

    array = asarray


ref.jnp = _JnpForDatasets(ref.jnp)


def np_(t):
    return t.detach().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)


def seeded_nu(n, seed):
    """nu(coords_chunk, omega): symmetric n x n per point, deterministic in (coords, omega)."""

    def nu(coords, omega):
        c = torch.as_tensor(np.asarray(coords), dtype=F64)
        g = torch.Generator().manual_seed(seed)
        basis = torch.randn(6, n, n, generator=g, dtype=F64)
        basis = basis + basis.transpose(1, 2)
        feats = torch.stack([torch.ones(len(c), dtype=F64), torch.cos(c[:, 0]), torch.sin(c[:, 1]), c[:, 2] / 6.0,
                             torch.exp(-float(omega) * (c ** 2).sum(1) / 20.0), torch.cos(c.sum(1) * (1.0 + float(omega)))], dim=1)
        return torch.einsum("rk,kab->rab", feats, basis)

    return nu


def ref_molecule(mol, name=None, with_chi=True, energy=None):
    g2 = {2: J(mol["grad_n_ao2"])} if "grad_n_ao2" in mol else None
    return molmod.Molecule(
        grid=molmod.Grid(J(mol["coords"]), J(mol["weights"])), atom_index=J(torch.tensor([8, 1, 1])), nuclear_pos=J(torch.arange(9, dtype=F64).reshape(3, 3)),
        ao=J(mol["ao"]), grad_ao=J(mol["grad_ao"]), grad_n_ao=g2, rdm1=J(mol["rdm1"]), nuclear_repulsion=J(mol["nuclear_repulsion"]),
        h1e=J(mol["h1e"]), vj=J(torch.stack([mol["h1e"], 2.0 * mol["h1e"]])), mo_coeff=J(mol["mo_coeff"]), mo_occ=J(mol["mo_occ"]), mo_energy=J(mol["mo_energy"]),
        s1e=J(mol["s1e"]), omegas=(J(mol["omegas"]) if with_chi else None), chi=(J(mol["chi"]) if with_chi else None),
        rep_tensor=J(mol["rep_tensor"]), energy=energy, name=([ord(c) for c in name] if name else None),
        basis=[ord(c) for c in "def2-tzvp"], spin=0, charge=0, scf_iteration=50,
    )


def molecule_fields(m):
    out = {}
    for k, v in m.to_dict().items():
        if v is None:
            out[f"{k}#none"] = np.zeros((), dtype=np.int8)
        elif isinstance(v, dict):
            for kk, vv in v.items():
                out[f"{k}.{kk}"] = np_(vv)
        elif isinstance(v, str):
            out[f"{k}#str"] = np.frombuffer(v.encode(), dtype=np.uint8)
        else:
            out[k] = np_(v) if isinstance(v, torch.Tensor) else np.asarray(v)
    return out


def main():
    # ---- 1. generate_chi_tensor ------------------------------------------------------------------------------
    d = {}
    for tag, (N, n, seed, chunk) in {"a": (203, 10, 1984, 37), "b": (97, 7, 1993, 1024)}.items():
        mol = synthetic_molecule(N, n, seed=seed, symmetric_rdm1=(tag == "a"), mask_frac=0.0, with_eri=False)
        omegas = [0.0, 0.4]
        provider = seeded_nu(n, seed + 7)

        def nu_chunk(mol_, coords, omega, chunk_size=1000):
            if omega < 0:
                raise ValueError("Range-separated parameter omega must be non-negative!")
            for i in range(0, len(coords), chunk_size):
                j = min(i + chunk_size, len(coords))
                yield i, j, J(provider(coords[i:j], omega))

        ref._nu_chunk = nu_chunk
        chi = ref.generate_chi_tensor(J(mol["rdm1"]), J(mol["ao"]), J(mol["coords"]), None, omegas, chunk_size=chunk)
        d[f"{tag}_ao"], d[f"{tag}_rdm1"], d[f"{tag}_coords"] = np_(mol["ao"]), np_(mol["rdm1"]), np_(mol["coords"])
        d[f"{tag}_omegas"], d[f"{tag}_chunk"], d[f"{tag}_nu_seed"] = np.array(omegas), np.array(chunk), np.array(seed + 7)
        d[f"{tag}_nu"] = np.stack([np_(provider(mol["coords"], o)) for o in omegas])
        d[f"{tag}_out_chi"] = np_(chi)
    np.savez_compressed(HERE / "io_chi.npz", **d)

    # ---- 2. saver: the tree -------------------------------------------------------------------------------------
    mols = {
        "water": synthetic_molecule(61, 5, n_omega=2, seed=1984, mask_frac=0.0),
        "anon": synthetic_molecule(47, 4, n_omega=2, seed=1993, mask_frac=0.0),
        "r1": synthetic_molecule(33, 3, n_omega=2, seed=7, mask_frac=0.0),
        "p1": synthetic_molecule(29, 3, n_omega=2, seed=8, mask_frac=0.0),
    }
    inputs = {}
    for k, m in mols.items():
        for f, v in m.items():
            inputs[f"{k}.{f}"] = np_(v)
    m_water = ref_molecule(mols["water"], name="water", energy=-76.4)
    m_anon = ref_molecule(mols["anon"], name=None, with_chi=False)
    m_r1 = ref_molecule(mols["r1"], name="r1", energy=-1.1)
    m_p1 = ref_molecule(mols["p1"], name="p1", energy=-0.5)
    reaction = molmod.make_reaction([m_r1], [m_p1, m_p1], [1], [1, 1], energy=0.1, name="diss")
    ref.saver("golden_io", reactions=[reaction], molecules=[m_water, m_anon])
    tree = fake_h5py.flatten(fake_h5py.FILES["golden_io.hdf5"])
    np.savez_compressed(HERE / "io_tree.npz", **tree)
    np.savez_compressed(HERE / "io_inputs.npz", **inputs)

    # ---- 3. loader ------------------------------------------------------------------------------------------------
    out = {}
    cases = {"train_all": dict(training=True, config_omegas=None), "eval_all": dict(training=False, config_omegas=None),
             "train_sel": dict(training=True, config_omegas=[0.4]), "train_nochi": dict(training=True, config_omegas=[])}
    # upstream's reaction branch looks `omegas` up in the REACTION group (pyscf.py:548) and raises KeyError whenever
    # config_omegas is a non-empty list, so the selection case reads a molecules-only file
    ref.saver("golden_io_mols", molecules=[m_water, m_anon])
    for case, kw in cases.items():
        fname = "golden_io_mols" if case == "train_sel" else "golden_io"
        for idx, (kind, obj) in enumerate(ref.loader(fname, randomize=False, **kw)):
            if kind == "molecule":
                for k, v in molecule_fields(obj).items():
                    out[f"{case}/{idx}/molecule/{k}"] = v
            else:
                out[f"{case}/{idx}/reaction/energy"] = np.asarray(float(obj.energy))
                out[f"{case}/{idx}/reaction/reactant_numbers"] = np.asarray([int(x) for x in obj.reactant_numbers])
                out[f"{case}/{idx}/reaction/product_numbers"] = np.asarray([int(x) for x in obj.product_numbers])
                if obj.name is not None:
                    out[f"{case}/{idx}/reaction/name"] = np_(obj.name)
                for role, ms in (("reactants", obj.reactants), ("products", obj.products)):
                    for j, m in enumerate(ms):
                        for k, v in molecule_fields(m).items():
                            out[f"{case}/{idx}/reaction/{role}/{j}/{k}"] = v
    np.savez_compressed(HERE / "io_loaded.npz", **out)
    print("tree keys:", len(tree), " loaded keys:", len(out))


if __name__ == "__main__":
    main()
