"""Row f4 on the CPU: the oracle's chi restatement and the saver / loader tree against the golden data produced by the
reference's own grad_dft/interface/pyscf.py (tests/golden/make_golden_io.py)."""
from pathlib import Path

import numpy as np
import pytest
import torch

import oracle
from graddft_b200 import interface
from graddft_b200.molecule import Grid, Molecule

G = Path(__file__).resolve().parent / "golden"
F64 = torch.float64


def seeded_nu(n, seed):
    """the nu provider of make_golden_io.py (same expression, same seed)"""

    def nu(coords, omega):
        c = torch.as_tensor(np.asarray(coords.cpu() if isinstance(coords, torch.Tensor) else coords), dtype=F64)
        g = torch.Generator().manual_seed(seed)
        basis = torch.randn(6, n, n, generator=g, dtype=F64)
        basis = basis + basis.transpose(1, 2)
        feats = torch.stack([torch.ones(len(c), dtype=F64), torch.cos(c[:, 0]), torch.sin(c[:, 1]), c[:, 2] / 6.0,
                             torch.exp(-float(omega) * (c ** 2).sum(1) / 20.0), torch.cos(c.sum(1) * (1.0 + float(omega)))], dim=1)
        return torch.einsum("rk,kab->rab", feats, basis)

    return nu


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_chi_matches_reference(tag):
    z = np.load(G / "io_chi.npz")
    t = lambda k: torch.from_numpy(z[f"{tag}_{k}"])  # noqa: E731
    n = t("ao").shape[1]
    provider = seeded_nu(n, int(z[f"{tag}_nu_seed"]))
    for w, omega in enumerate(z[f"{tag}_omegas"]):  # the provider reproduces the stored nu
        assert torch.allclose(provider(t("coords"), float(omega)), t("nu")[w], rtol=0, atol=1e-13)
    chi = oracle.generate_chi_tensor(t("rdm1"), t("ao"), t("coords"), provider, [float(o) for o in z[f"{tag}_omegas"]], int(z[f"{tag}_chunk"]))
    ref = t("out_chi")
    assert chi.shape == ref.shape
    assert float((chi - ref).abs().max()) <= 1e-12 * float(ref.abs().max())
    with pytest.raises(ValueError):
        oracle.generate_chi_tensor(t("rdm1"), t("ao"), t("coords"), provider, [-0.1])


def _molecule(fields, name=None, with_chi=True, energy=None):
    f = {k: torch.from_numpy(np.asarray(v)) for k, v in fields.items()}
    return Molecule(
        grid=Grid(f["coords"], f["weights"]), atom_index=torch.tensor([8, 1, 1]), nuclear_pos=torch.arange(9, dtype=F64).reshape(3, 3),
        ao=f["ao"], grad_ao=f["grad_ao"], grad_n_ao={2: f["grad_n_ao2"]}, rdm1=f["rdm1"], nuclear_repulsion=f["nuclear_repulsion"],
        h1e=f["h1e"], vj=torch.stack([f["h1e"], 2.0 * f["h1e"]]), mo_coeff=f["mo_coeff"], mo_occ=f["mo_occ"], mo_energy=f["mo_energy"],
        s1e=f["s1e"], omegas=(f["omegas"] if with_chi else None), chi=(f["chi"] if with_chi else None), rep_tensor=f["rep_tensor"],
        energy=energy, name=([ord(c) for c in name] if name else None), basis=[ord(c) for c in "def2-tzvp"], spin=0, charge=0,
        scf_iteration=50,
    )


@pytest.fixture()
def golden_molecules():
    z = np.load(G / "io_inputs.npz")
    per = {}
    for k in z.files:
        mol, field = k.split(".", 1)
        per.setdefault(mol, {})[field] = z[k]
    water = _molecule(per["water"], "water", energy=-76.4)
    anon = _molecule(per["anon"], None, with_chi=False)
    r1 = _molecule(per["r1"], "r1", energy=-1.1)
    p1 = _molecule(per["p1"], "p1", energy=-0.5)
    reaction = interface.make_reaction([r1], [p1, p1], [1], [1, 1], energy=0.1, name="diss")
    return water, anon, reaction


def test_saver_writes_the_reference_tree(tmp_path, golden_molecules, monkeypatch):
    monkeypatch.setattr(interface, "_h5py", None)  # the Archive backend (h5py is absent here anyway)
    water, anon, reaction = golden_molecules
    path = interface.saver(str(tmp_path / "data.hdf5"), reactions=[reaction], molecules=[water, anon])
    assert path.endswith("data.npz")
    ours = dict(np.load(path))
    ref = np.load(G / "io_tree.npz")
    assert sorted(ours) == sorted(ref.files)
    for k in ref.files:
        a, b = ours[k], ref[k]
        assert a.shape == b.shape, k
        assert a.dtype.kind == b.dtype.kind, (k, a.dtype, b.dtype)
        assert np.array_equal(a, b), k
    # appending to an existing file: a second molecule group with a fresh index, an existing name is refused (h5py does)
    with pytest.raises(ValueError):
        interface.saver(str(tmp_path / "data"), molecules=[water])


def _fields(m):
    out = {}
    for k, v in m.to_dict().items():
        if v is None:
            out[f"{k}#none"] = None
        elif isinstance(v, dict):
            for kk, vv in v.items():
                out[f"{k}.{kk}"] = vv
        elif isinstance(v, str):
            out[f"{k}#str"] = np.frombuffer(v.encode(), dtype=np.uint8)
        else:
            out[k] = v
    return out


def _same(ours, ref, key):
    if ours is None:
        assert ref.shape == () and ref.dtype == np.int8, key
        return
    a = ours.numpy() if isinstance(ours, torch.Tensor) else np.asarray(ours)
    assert a.shape == ref.shape, (key, a.shape, ref.shape)
    assert a.dtype.kind == ref.dtype.kind, (key, a.dtype, ref.dtype)
    assert np.array_equal(a, ref), key


@pytest.mark.parametrize("case,kw", [
    ("train_all", dict(training=True, config_omegas=None)), ("eval_all", dict(training=False, config_omegas=None)),
    ("train_sel", dict(training=True, config_omegas=[0.4])), ("train_nochi", dict(training=True, config_omegas=[])),
])
def test_loader_yields_what_the_reference_yields(tmp_path, golden_molecules, monkeypatch, case, kw):
    monkeypatch.setattr(interface, "_h5py", None)
    water, anon, reaction = golden_molecules
    if case == "train_sel":
        interface.saver(str(tmp_path / "d"), molecules=[water, anon])
    else:
        interface.saver(str(tmp_path / "d"), reactions=[reaction], molecules=[water, anon])
    ref = np.load(G / "io_loaded.npz")
    want = {k[len(case) + 1:]: ref[k] for k in ref.files if k.startswith(case + "/")}
    got = {}
    for idx, (kind, obj) in enumerate(interface.loader(str(tmp_path / "d"), randomize=False, **kw)):
        if kind == "molecule":
            for k, v in _fields(obj).items():
                got[f"{idx}/molecule/{k}"] = v
        else:
            got[f"{idx}/reaction/energy"] = np.asarray(float(obj.energy))
            got[f"{idx}/reaction/reactant_numbers"] = np.asarray([int(x) for x in obj.reactant_numbers])
            got[f"{idx}/reaction/product_numbers"] = np.asarray([int(x) for x in obj.product_numbers])
            if obj.name is not None:
                got[f"{idx}/reaction/name"] = obj.name
            for role, ms in (("reactants", obj.reactants), ("products", obj.products)):
                for j, m in enumerate(ms):
                    for k, v in _fields(m).items():
                        got[f"{idx}/reaction/{role}/{j}/{k}"] = v
    assert sorted(got) == sorted(want)
    for k in want:
        _same(got[k], want[k], k)


def test_loader_selects_omegas_inside_reactions(tmp_path, golden_molecules, monkeypatch):
    """upstream raises KeyError here (it looks `omegas` up in the reaction group); the molecule's own list is used"""
    monkeypatch.setattr(interface, "_h5py", None)
    water, anon, reaction = golden_molecules
    interface.saver(str(tmp_path / "d"), reactions=[reaction])
    ((kind, r),) = list(interface.loader(str(tmp_path / "d"), config_omegas=[0.4, 0.0]))
    assert kind == "reaction"
    full = reaction.reactants[0].chi
    assert torch.equal(r.reactants[0].chi, torch.stack([full[:, 1], full[:, 0]], dim=1))
    with pytest.raises(AssertionError):
        list(interface.loader(str(tmp_path / "d"), config_omegas=[0.7]))


def test_generate_chi_needs_cuda():
    from graddft_b200._lib import GdftError
    with pytest.raises(GdftError):
        interface.generate_chi_tensor(torch.zeros(2, 3, 3, dtype=F64), torch.zeros(5, 3, dtype=F64), torch.zeros(5, 3, dtype=F64),
                                      lambda c, o: torch.zeros(len(c), 3, 3, dtype=F64), [0.0])


def test_archive_append_and_attrs(tmp_path):
    """Archive (the h5py stand-in): groups / datasets / attributes / string datasets survive close + reopen, "a" appends,
    duplicate names are refused, iteration is name-ordered (what h5py does)."""
    path = str(tmp_path / "x.npz")
    with interface.Archive(path, "a") as f:
        g = f.create_group("molecule_b_1")
        g.create_dataset("ao", data=np.arange(6.0).reshape(2, 3))
        g.create_dataset("name", data="H2O")
        g.attrs["type"] = "reactant"
        f.create_group("empty_group")
    with interface.Archive(path, "a") as f:
        f.create_group("molecule_a_0")["energy"] = -1.5
        with pytest.raises(ValueError):
            f.create_group("molecule_b_1")
    with interface.Archive(path, "r") as f:
        assert [k for k, _ in f.items()] == ["empty_group", "molecule_a_0", "molecule_b_1"]
        g = f["molecule_b_1"]
        assert np.array_equal(np.asarray(g["ao"]), np.arange(6.0).reshape(2, 3))
        assert g["name"][()] == b"H2O" and str(g["name"][()]) == "b'H2O'"   # h5py hands strings back as bytes
        assert g.attrs["type"] == "reactant"
        assert float(f["molecule_a_0/energy"][()]) == -1.5
    with pytest.raises(FileNotFoundError):
        interface.Archive(str(tmp_path / "missing.npz"), "r")


def test_loader_shuffles_only_when_training(tmp_path, golden_molecules, monkeypatch):
    monkeypatch.setattr(interface, "_h5py", None)
    water, anon, reaction = golden_molecules
    interface.saver(str(tmp_path / "d"), reactions=[reaction], molecules=[water, anon])
    order = [k for k, _ in interface.loader(str(tmp_path / "d"), randomize=True, training=False)]
    assert order == ["molecule", "molecule", "reaction"]  # name order: no shuffle outside training (pyscf.py:464-465)
    import random
    random.seed(3)
    seen = {tuple(k for k, _ in interface.loader(str(tmp_path / "d"), randomize=True, training=True)) for _ in range(12)}
    assert len(seen) > 1
