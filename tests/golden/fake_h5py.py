"""An in-memory stand-in for the subset of h5py that grad_dft/interface/pyscf.py's saver / loader use, so the
reference's OWN saver and loader can be executed here (h5py is not installed) and the tree they write / the molecules
they read back can be committed as golden data.  TEST INFRASTRUCTURE for tests/golden/make_golden_io.py only.

h5py semantics reproduced: string datasets read back as bytes; `dataset[()]` gives the scalar / array; groups iterate
in name order; creating an existing name raises ValueError; attrs is a dict; files opened with "a" persist across
opens of the same path (kept in FILES)."""
import numpy as np

FILES = {}


class Dataset:
    def __init__(self, value):
        self.value = value

    def __getitem__(self, key):
        if isinstance(self.value, bytes):
            return self.value
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

    def __index__(self):
        return int(np.asarray(self.value))

    @property
    def shape(self):
        return () if isinstance(self.value, bytes) else np.asarray(self.value).shape


class Group:
    def __init__(self):
        self.children = {}
        self.attrs = {}

    def create_group(self, name):
        if name in self.children:
            raise ValueError(f"Unable to create group (name already exists): {name}")
        g = self.children[name] = Group()
        return g

    def create_dataset(self, name, shape=None, chunks=None, dtype=None, data=None):
        if name in self.children:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        if isinstance(data, str):
            v = data.encode()
        else:
            if hasattr(data, "detach"):
                data = data.detach().numpy()
            v = np.array(data)
        d = self.children[name] = Dataset(v)
        return d

    def __setitem__(self, name, value):
        self.create_dataset(name, data=value)

    def __getitem__(self, name):
        return self.children[name]

    def items(self):
        return [(k, self.children[k]) for k in sorted(self.children)]


class File(Group):
    def __new__(cls, path, mode="r"):
        if mode == "r":
            return FILES[path]
        if path not in FILES:
            f = super().__new__(cls)
            Group.__init__(f)
            FILES[path] = f
        return FILES[path]

    def __init__(self, path, mode="r"):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def flatten(group, prefix=""):
    """{path: array}: datasets under their path, string datasets under "<path>#s" (uint8), attributes under "@<name>"."""
    out = {}
    if prefix and not group.children and not group.attrs:
        out[prefix + "#group"] = np.zeros((), dtype=np.int8)
    for k, v in group.attrs.items():
        out[f"{prefix}@{k}"] = np.array(v)
    for k, v in group.children.items():
        if isinstance(v, Group):
            out.update(flatten(v, f"{prefix}{k}/"))
        elif isinstance(v.value, bytes):
            out[f"{prefix}{k}#s"] = np.frombuffer(v.value, dtype=np.uint8)
        else:
            out[f"{prefix}{k}"] = v.value
    return out
