"""Row f4: the chi-generation tail and the on-disk molecule format of
grad_dft/interface/pyscf.py, the two things on either side of the hot path that are not PySCF's.

* `generate_chi_tensor` (pyscf.py:1062-1124): for every omega and every chunk of grid points, nu[chunk, n, n] (the
  screened-Coulomb integrals of `_nu_chunk`, grad_dft/external/_hf_density.py:69-103 -- libcint, out of path) is
  contracted with ao and rdm1 by ONE streaming kernel (`gdft_chi_contract`) that writes straight into chi[N, W, 2, n];
  host-resident chunks are staged through two pinned buffers on a copy stream so the upload of chunk k+1 overlaps the
  contraction of chunk k.  Where the reference takes a PySCF `Mole`, this takes the nu provider (any callable
  `nu(coords_chunk, omega) -> [chunk, n, n]`); a `Mole` is accepted when PySCF is importable.
* `saver` / `loader` / `save_molecule_data` (pyscf.py:330-609): the same group / dataset / attribute tree, written
  through h5py when it is importable and otherwise through `Archive`, a single-file stand-in (`<fname>.npz`, keys are
  the HDF5 paths) exposing the subset of the h5py API the tree needs.  Upstream quirks are kept:
  fields that are None are not written, `name` / `basis` come back as the characters of str(bytes) ("b'...'"), and
  every other dataset -- the integer `atom_index` included -- is returned as float64.
"""
from __future__ import annotations

import os
from itertools import chain
from random import shuffle
from typing import Any, Callable, Dict, Iterator, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops
from .molecule import Grid, Molecule, Reaction

try:  # pragma: no cover - not installed in this image
    import h5py as _h5py
except Exception:  # noqa: BLE001
    _h5py = None

F64 = torch.float64
Array = torch.Tensor


# ---------------------------------------------------------------------------------------------------------
# chi generation
# ---------------------------------------------------------------------------------------------------------
def _pyscf_nu(mol) -> Callable:
    """nu(coords, omega) from a PySCF Mole: int1e_grids_sph under with_range_coulomb (external/_hf_density.py:34-46)."""

    def nu(coords, omega):
        with mol.with_range_coulomb(omega=float(omega)):
            return mol.intor("int1e_grids_sph", hermi=1, grids=np.asarray(coords))

    return nu


_STAGE_THREADS = max(1, min(8, (os.cpu_count() or 2) // 2))
_POOL = None


def _stage_pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor

        _POOL = ThreadPoolExecutor(max_workers=_STAGE_THREADS, thread_name_prefix="gdft-stage")
    return _POOL


def _copy_rows(dst: Array, src: Array, a: int, b: int) -> None:
    """dst[a:b] <- src[a:b] on the host.  Contiguous float64 slices go through memmove from a foreign-function call, which
    runs without the interpreter lock, so the pool's threads really copy in parallel."""
    if src.is_contiguous() and dst.is_contiguous() and src.dtype == dst.dtype and b > a:
        import ctypes

        row = src.stride(0) * src.element_size()
        ctypes.memmove(dst.data_ptr() + a * row, src.data_ptr() + a * row, (b - a) * row)
    else:
        dst[a:b].copy_(src[a:b])


class _Uploader:
    """Double-buffered host -> device staging of nu chunks: pinned buffer + copy stream, the compute stream waits on the
    copy's event and the copy stream waits until the kernel that read the device buffer two chunks ago has finished."""

    def __init__(self, device, chunk: int, n: int):
        self.device = device
        self.shape = (chunk, n)
        self.copy_stream = torch.cuda.Stream(device=device)
        # the freshly allocated device buffers below may be blocks the caching allocator has just taken back from
        # kernels still running on the compute stream: the copy stream must not write into them before those finish
        self.copy_stream.wait_stream(torch.cuda.current_stream(device))
        self.inflight = None  # copy event of a chunk uploaded straight from the provider's own page-locked buffer
        self.pinned = [torch.empty((chunk, n, n), dtype=F64).pin_memory() for _ in range(2)]
        self.dev = [torch.empty((chunk, n, n), dtype=F64, device=device) for _ in range(2)]
        self.free = [None, None]  # event recorded after the kernel that consumed dev[i]
        self.keep = [None, None]
        self.k = 0
        self.h2d_bytes = 0

    def stage(self, host_chunk) -> Array:
        i = self.k & 1
        self.k += 1
        src = torch.as_tensor(np.asarray(host_chunk) if not isinstance(host_chunk, torch.Tensor) else host_chunk, dtype=F64)
        m = src.shape[0]
        if self.free[i] is not None:
            self.free[i].synchronize()  # the pinned buffer is also reused: the host must not overwrite it early
        if src.is_pinned() and src.is_contiguous():
            staged = src                # the provider already wrote into page-locked memory: no staging copy
            self.keep[i] = src          # alive until the copy has been consumed
            with torch.cuda.stream(self.copy_stream):
                self.dev[i][:m].copy_(staged, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record()
        else:
            # pageable source: one host thread copies at ~9 GB/s, a quarter of what the link takes.  The chunk is cut into
            # row slices staged by a small pool of threads (the copy releases the interpreter lock), and each slice's upload
            # is enqueued as soon as it is staged, so staging and upload overlap inside the chunk as well
            staged = self.pinned[i][:m]
            parts = max(1, min(_STAGE_THREADS, m))
            bounds = [(m * k // parts, m * (k + 1) // parts) for k in range(parts)]
            futs = [_stage_pool().submit(_copy_rows, staged, src, a, b) for a, b in bounds]
            with torch.cuda.stream(self.copy_stream):
                for (a, b), fut in zip(bounds, futs):
                    fut.result()
                    self.dev[i][a:b].copy_(staged[a:b], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record()
        torch.cuda.current_stream(self.device).wait_event(ready)
        self.inflight = ready if staged is src else None
        self.h2d_bytes += m * src.shape[1] * src.shape[2] * 8
        return self.dev[i][:m], i

    def before_next_chunk(self) -> None:
        """Called before the provider is asked for the next chunk: a provider that hands over page-locked memory may
        reuse that one buffer for every chunk, so the upload reading it must have finished (the kernel consuming the
        device copy keeps running meanwhile)."""
        if self.inflight is not None:
            self.inflight.synchronize()
            self.inflight = None

    def consumed(self, i: int) -> None:
        ev = torch.cuda.Event()
        ev.record()
        self.free[i] = ev


_UPLOADERS: Dict[Any, "_Uploader"] = {}


def _uploader_for(device, chunk: int, n: int) -> "_Uploader":
    """The staging buffers of (device, chunk, n), kept between calls: page-locking 2 x chunk x n^2 doubles costs far more
    than a whole chi build (1.1 GB at the benzene shape: ~0.2 s), and an SCF loop regenerates chi every cycle
    (grad_dft/evaluate.py:511-522).  One entry per device; a different shape replaces it."""
    key = torch.device(device).index
    up = _UPLOADERS.get(key)
    if up is None or up.shape != (chunk, n):
        up = _UPLOADERS[key] = _Uploader(device, chunk, n)
    else:
        up.copy_stream.wait_stream(torch.cuda.current_stream(device))
        up.inflight = None
    return up


def generate_chi_tensor(rdm1: Array, ao: Array, grid_coords: Array, mol: Any, omegas: Sequence[float], chunk_size: Optional[int] = 1024,
                        precision: Any = None, *args, **kwargs) -> Array:
    """grad_dft/interface/pyscf.py:1062-1124.  chi[r, w, s, a] = sum_{b,d} rdm1[s,b,d] ao[r,b] nu_w[r,d,a], shape
    (n_grid, n_omega, 2, n_orbitals); an empty omega list gives an empty tensor.  `mol` is the nu provider (see module
    docstring).  Raises ValueError for a negative omega (external/_hf_density.py:92-93)."""
    del precision, args, kwargs  # float64 throughout
    if not ao.is_cuda:
        raise ops._lib.GdftError("graddft_b200 kernels need CUDA tensors (there is no CPU path)")
    nu_fn = mol if callable(mol) else _pyscf_nu(mol)
    omegas = [float(o) for o in (omegas.tolist() if isinstance(omegas, torch.Tensor) else omegas)]
    if not omegas:
        return torch.zeros((0,), dtype=F64, device=ao.device)
    if any(o < 0 for o in omegas):
        raise ValueError("Range-separated parameter omega must be non-negative!")
    N, n = int(ao.shape[0]), int(ao.shape[1])
    if chunk_size is None:
        chunk_size = N
    ao = ao.contiguous()
    rdm1 = rdm1.detach().contiguous()
    chi = torch.empty((N, len(omegas), 2, n), dtype=F64, device=ao.device)
    up = None
    for w, omega in enumerate(omegas):
        for start in range(0, N, chunk_size):
            end = min(start + chunk_size, N)
            if up is not None:
                up.before_next_chunk()
            nu = nu_fn(grid_coords[start:end], omega)
            slot = None
            if not (isinstance(nu, torch.Tensor) and nu.is_cuda):
                if up is None:
                    up = _uploader_for(ao.device, min(chunk_size, N), n)
                nu, slot = up.stage(nu)
            if tuple(nu.shape) != (end - start, n, n):
                raise TypeError(f"nu chunk has shape {tuple(nu.shape)}, expected {(end - start, n, n)}")
            ops.chi_contract_(chi, start, w, ao, rdm1, nu)
            if slot is not None:
                up.consumed(slot)
    return chi


# ---------------------------------------------------------------------------------------------------------
# Archive: the h5py subset the tree needs, on one .npz file
# ---------------------------------------------------------------------------------------------------------
class _Dataset:
    def __init__(self, value):
        self.value = value  # np.ndarray, or bytes for a string dataset (what h5py returns for one)

    def __getitem__(self, key):
        if isinstance(key, tuple) and key == ():
            return self.value if isinstance(self.value, bytes) else (self.value[()] if self.value.shape == () else self.value)
        return self.value[key]

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self.value)
        return a.astype(dtype) if dtype is not None else a

    def __iter__(self):
        return iter(np.asarray(self.value))

    def __float__(self):
        return float(np.asarray(self.value))

    def __int__(self):
        return int(np.asarray(self.value))

    @property
    def shape(self):
        return () if isinstance(self.value, bytes) else self.value.shape


class _Group:
    def __init__(self):
        self.children: Dict[str, Union["_Group", _Dataset]] = {}
        self.attrs: Dict[str, Any] = {}

    def create_group(self, name: str) -> "_Group":
        if name in self.children:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        g = self.children[name] = _Group()
        return g

    def create_dataset(self, name: str, data=None, **_ignored) -> _Dataset:
        if name in self.children:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        if isinstance(data, str):
            v = data.encode()
        elif isinstance(data, bytes):
            v = data
        else:
            v = np.array(data.detach().cpu().numpy() if isinstance(data, torch.Tensor) else data)
        d = self.children[name] = _Dataset(v)
        return d

    def __setitem__(self, name, value):
        self.create_dataset(name, data=value)

    def __getitem__(self, name):
        node = self
        for part in name.split("/"):
            node = node.children[part]
        return node

    def __contains__(self, name):
        return name in self.children

    def items(self):
        return [(k, self.children[k]) for k in sorted(self.children)]  # h5py iterates groups in name order

    def keys(self):
        return sorted(self.children)


class Archive(_Group):
    """`h5py.File(path, mode)` stand-in for the molecule tree: modes "r", "a" and "w"; the whole tree lives in memory and
    is written as one .npz on close.  Keys: "<group>/<dataset>" for arrays, "<...>#s" for string datasets,
    "<group>/@<attr>" for attributes."""

    def __init__(self, path: str, mode: str = "r"):
        super().__init__()
        self.path, self.mode = path, mode
        if mode in ("r", "a") and os.path.exists(path):
            with np.load(path, allow_pickle=False) as z:
                for key in z.files:
                    self._insert(key, z[key])
        elif mode == "r":
            raise FileNotFoundError(path)

    def _insert(self, key: str, arr: np.ndarray) -> None:
        parts = key.split("/")
        node = self
        for part in parts[:-1]:
            node = node.children.setdefault(part, _Group())
        leaf = parts[-1]
        if leaf.startswith("@"):
            node.attrs[leaf[1:]] = arr.item() if arr.shape == () else arr
        elif leaf.endswith("#s"):
            node.children[leaf[:-2]] = _Dataset(bytes(arr.tobytes()))
        elif leaf == "#group":
            pass
        else:
            node.children[leaf] = _Dataset(arr)

    def flatten(self) -> Dict[str, np.ndarray]:
        out: Dict[str, np.ndarray] = {}

        def walk(g: _Group, prefix: str):
            if prefix and not g.children and not g.attrs:
                out[prefix + "#group"] = np.zeros((), dtype=np.int8)
            for k, v in g.attrs.items():
                out[f"{prefix}@{k}"] = np.array(v)
            for k, v in g.children.items():
                if isinstance(v, _Group):
                    walk(v, f"{prefix}{k}/")
                elif isinstance(v.value, bytes):
                    out[f"{prefix}{k}#s"] = np.frombuffer(v.value, dtype=np.uint8)
                else:
                    out[f"{prefix}{k}"] = v.value

        walk(self, "")
        return out

    def close(self) -> None:
        if self.mode in ("a", "w"):
            tmp = self.path + ".tmp.npz"
            np.savez(tmp, **self.flatten())
            os.replace(tmp, self.path)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if exc[0] is None:
            self.close()
        return False


def _open(fname: str, mode: str):
    """(file object, path): `<fname>.hdf5` through h5py when it is importable, else `<fname>.npz` through Archive."""
    fname = fname.replace(".hdf5", "").replace(".h5", "").replace(".npz", "")
    if _h5py is not None and (mode != "r" or os.path.exists(f"{fname}.hdf5")):
        return _h5py.File(os.path.normpath(f"{fname}.hdf5"), mode)
    return Archive(os.path.normpath(f"{fname}.npz"), mode)


# ---------------------------------------------------------------------------------------------------------
# saver / loader
# ---------------------------------------------------------------------------------------------------------
def make_reaction(reactants, products, reactant_numbers=None, product_numbers=None, energy=None, name=None) -> Reaction:
    """grad_dft/molecule.py:905-958."""

    def canon(molecules, numbers):
        if isinstance(molecules, Molecule):
            molecules = (molecules,)
        if numbers is None:
            numbers = (1,) * len(molecules)
        if len(numbers) != len(molecules):
            raise ValueError("the number of multiplicities must match the number of molecules")
        return molecules, numbers

    reactants, reactant_numbers = canon(reactants, reactant_numbers)
    products, product_numbers = canon(products, product_numbers)
    return Reaction(reactants, products, reactant_numbers, product_numbers, energy, name)


def _np(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return x


def _chars(codes) -> str:
    return "".join(chr(int(c)) for c in (codes.tolist() if isinstance(codes, (torch.Tensor, np.ndarray)) else codes))


def save_molecule_data(mol_group, molecule: Molecule) -> None:
    """grad_dft/interface/pyscf.py:572-589: one dataset per Molecule field (grid coords/weights first); None fields are
    not written; name/basis as strings; grad_n_ao as a sub-group keyed by the derivative order."""
    for name, data in molecule.to_dict().items():
        if data is None:
            continue
        if name in ("name", "basis"):
            mol_group.create_dataset(name, data=data if isinstance(data, str) else _chars(data))
        elif name == "grad_n_ao":
            g = mol_group.create_group(name)
            for k, v in data.items():
                g.create_dataset(f"{k}", data=_np(v))
        else:
            mol_group.create_dataset(name, data=_np(data) if isinstance(data, (torch.Tensor, np.ndarray)) else np.asarray(data))


def _group_name(prefix: str, name, index: int) -> str:
    if name is not None:
        return f"{prefix}_{name if isinstance(name, str) else _chars(name)}_{index}"
    return f"{prefix}_{index}"


def saver(fname: str, reactions: Union[Reaction, Sequence[Reaction]] = (), molecules: Union[Molecule, Sequence[Molecule]] = ()) -> str:
    """grad_dft/interface/pyscf.py:330-426: appends the reactions and molecules to `<fname>.hdf5` (h5py) or `<fname>.npz`
    (Archive).  Returns the path written."""
    if isinstance(molecules, Molecule):
        molecules = (molecules,)
    if isinstance(reactions, Reaction):
        reactions = (reactions,)
    f = _open(fname, "a")
    with f as file:
        for i, reaction in enumerate(reactions):
            # upstream interpolates reaction.name as is (a str); molecule names are integer code points
            react = file.create_group(f"reaction_{reaction.name}_{i}" if reaction.name else f"reaction_{i}")
            react["energy"] = reaction.energy
            for j, molecule in enumerate(chain(reaction.reactants, reaction.products)):
                mol_group = react.create_group(_group_name("molecule", molecule.name, j))
                save_molecule_data(mol_group, molecule)
                if j < len(reaction.reactants):
                    mol_group.attrs["type"] = "reactant"
                    mol_group["reactant_numbers"] = reaction.reactant_numbers[j]
                else:
                    mol_group.attrs["type"] = "product"
                    mol_group["product_numbers"] = reaction.product_numbers[j - len(reaction.reactant_numbers)]
        for j, molecule in enumerate(molecules):
            mol_group = file.create_group(_group_name("molecule", molecule.name, j))
            save_molecule_data(mol_group, molecule)
    return getattr(f, "path", None) or f"{fname}.hdf5"


def _t(value, device, dtype=F64) -> Array:
    return torch.as_tensor(np.asarray(value), dtype=dtype).to(device) if device is not None else torch.as_tensor(np.asarray(value), dtype=dtype)


def _read_molecule(group, omegas_holder, config_omegas, device, in_reaction: bool) -> Tuple[dict, Any]:
    """The per-field conversions of pyscf.py:468-498 (top-level molecules) and 529-560 (molecules of a reaction: no
    dtype coercion there except where upstream has one)."""
    args: Dict[str, Any] = {}
    for key, value in group.items():
        if key in ("reactant_numbers", "product_numbers"):
            continue
        if key in ("name", "basis"):
            args[key] = torch.tensor([ord(ch) for ch in str(value[()])], dtype=torch.int64)
        elif key == "energy":
            args[key] = torch.tensor(float(np.asarray(value[()])), dtype=F64)
        elif key in ("scf_iteration", "spin", "charge") and not in_reaction:
            args[key] = torch.tensor(int(np.asarray(value[()])), dtype=torch.int64)
        elif key == "grad_n_ao":
            args[key] = {int(k): _t(v, device) for k, v in value.items()}
        elif key == "chi":
            if config_omegas is None:
                args[key] = _t(value, device)
            elif list(config_omegas) == []:
                args[key] = None
            else:
                omegas = [float(o) for o in omegas_holder["omegas"]]
                missing = [o for o in config_omegas if float(o) not in omegas]
                assert not missing, f"chi tensors for omega list {config_omegas} were not all precomputed in the molecule"
                idx = [omegas.index(float(o)) for o in config_omegas]
                full = _t(value, device)
                args[key] = torch.stack([full[:, i] for i in idx], dim=1)
        else:
            a = np.asarray(value)
            args[key] = _t(a, device) if (not in_reaction or a.dtype.kind == "f") else _t(a, device, dtype=None)
    return args, group.attrs


def loader(fname: str, randomize: Optional[bool] = True, training: Optional[bool] = True,
           config_omegas: Optional[Sequence[float]] = None, device=None) -> Iterator[Tuple[str, Union[Molecule, Reaction]]]:
    """grad_dft/interface/pyscf.py:429-570: yields ("molecule", Molecule) / ("reaction", Reaction) for every top-level
    group, shuffled when `randomize and training`; `config_omegas` selects (and orders) the stored chi slices, [] drops
    chi.  Tensors land on `device` (default: host; pass "cuda" for kernels-ready molecules)."""
    with _open(fname, "r") as file:
        items = list(file.items())
        if randomize and training:
            shuffle(items)
        for grp_name, group in items:
            if "molecule" in grp_name:
                args, attrs = _read_molecule(group, group, config_omegas, device, in_reaction=False)
                if not training:
                    for key, value in attrs.items():
                        args[key] = str(value)
                grid = Grid(args.pop("coords"), args.pop("weights"))
                yield "molecule", _make_molecule(grid, args)
            if "reaction" in grp_name:
                reactants, products, reactant_numbers, product_numbers = [], [], [], []
                energy = torch.tensor(float(np.asarray(group["energy"][()])), dtype=F64)
                name = None if training else torch.tensor([ord(ch) for ch in str(grp_name.split("_")[1:])], dtype=torch.int64)
                for molecule_name, mgroup in group.items():
                    if molecule_name == "energy":
                        continue
                    # upstream reads `omegas` from the reaction group here (pyscf.py:548), a KeyError for every
                    # non-empty config_omegas; the molecule's own list is used instead
                    args, attrs = _read_molecule(mgroup, mgroup, config_omegas, device, in_reaction=True)
                    if not training:
                        args.setdefault("name", molecule_name.split("_")[1])  # overwritten by a stored name (pyscf.py:527-533)
                        for key, value in attrs.items():
                            if key != "type":
                                args[key] = value
                    grid = Grid(args.pop("coords"), args.pop("weights"))
                    mtype = attrs["type"]
                    mtype = mtype.decode() if isinstance(mtype, bytes) else str(mtype)
                    if mtype == "reactant":
                        reactants.append(_make_molecule(grid, args))
                        reactant_numbers.append(int(np.asarray(mgroup["reactant_numbers"][()])))
                    else:
                        products.append(_make_molecule(grid, args))
                        product_numbers.append(int(np.asarray(mgroup["product_numbers"][()])))
                yield "reaction", make_reaction(reactants, products, reactant_numbers, product_numbers, energy, name)


def _make_molecule(grid: Grid, args: dict) -> Molecule:
    """`Molecule(grid, **args)` (pyscf.py:506); positional fields that were None when saved were not written and come
    back as None here (upstream raises a TypeError for them instead)."""
    for f in ("atom_index", "nuclear_pos", "ao", "grad_ao", "grad_n_ao", "rdm1", "nuclear_repulsion", "h1e", "vj", "mo_coeff",
              "mo_occ", "mo_energy"):
        args.setdefault(f, None)
    return Molecule(grid, **args)
