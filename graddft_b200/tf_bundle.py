"""Reader for TensorFlow "tensor bundle" checkpoints (`variables/variables.index` + `variables.data-*`), without
TensorFlow: what `tf.saved_model.load(folder).variables` yields in grad_dft/functional.py:852-856, restated from the
published on-disk formats.

* `variables.index` is a LevelDB-format sorted string table (tensorflow/core/lib/io/table_format.txt): a 48-byte
  footer (two varint block handles, padding, magic 0xdb4775248b80fb57), an index block whose values are handles of
  the data blocks, and data blocks of prefix-compressed (shared, non_shared, value_len, key suffix, value) entries
  followed by a restart array.  Every block carries a 1-byte compression tag (0 = none, 1 = snappy) and a CRC.
* values are `BundleEntryProto` messages (tensorflow/core/protobuf/tensor_bundle.proto): dtype (1), shape (2),
  shard_id (3), offset (4), size (5), crc32c (6, masked) -- the key "" holds the `BundleHeaderProto`.
* tensor bytes are raw little-endian arrays at [offset, offset + size) of shard `shard_id`.

Host-side file parsing only (NumPy); nothing here is on the per-iteration path.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}


class BundleError(ValueError):
    pass


def _varint(b: bytes, p: int) -> Tuple[int, int]:
    r = s = 0
    while True:
        c = b[p]
        p += 1
        r |= (c & 0x7F) << s
        s += 7
        if c < 0x80:
            return r, p


def snappy_uncompress(b: bytes) -> bytes:
    """Raw snappy block format (format_description.txt): varint length, then literal / copy elements."""
    n, p = _varint(b, 0)
    out = bytearray()
    while p < len(b):
        t = b[p]
        p += 1
        kind = t & 3
        if kind == 0:  # literal
            ln = t >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(b[p:p + nb], "little")
                p += nb
            ln += 1
            out += b[p:p + ln]
            p += ln
            continue
        if kind == 1:
            ln, off = ((t >> 2) & 7) + 4, ((t >> 5) << 8) | b[p]
            p += 1
        elif kind == 2:
            ln, off = (t >> 2) + 1, int.from_bytes(b[p:p + 2], "little")
            p += 2
        else:
            ln, off = (t >> 2) + 1, int.from_bytes(b[p:p + 4], "little")
            p += 4
        if off == 0 or off > len(out):
            raise BundleError("corrupt snappy stream")
        for _ in range(ln):  # overlapping copies are legal (run-length)
            out.append(out[-off])
    if len(out) != n:
        raise BundleError("snappy length mismatch")
    return bytes(out)


_CRC_TABLE = None


def crc32c(data: bytes) -> int:
    """CRC-32C (Castagnoli), table-driven, vectorised over 4 KiB slabs is not needed at checkpoint sizes."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    c = 0xFFFFFFFF
    tab = _CRC_TABLE
    for x in data:
        c = tab[(c ^ x) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _unmask(m: int) -> int:
    """tensorflow/core/lib/hash/crc32c.h: Mask(crc) = rotr(crc, 15) + 0xa282ead8."""
    rot = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


def _block(buf: bytes, off: int, size: int, verify: bool):
    if off + size + 5 > len(buf):
        raise BundleError("block handle outside the file")
    body, tag = buf[off:off + size], buf[off + size]
    if verify:
        stored = _unmask(struct.unpack_from("<I", buf, off + size + 1)[0])
        if stored != crc32c(buf[off:off + size + 1]):
            raise BundleError("index block checksum mismatch")
    if tag == 1:
        body = snappy_uncompress(body)
    elif tag != 0:
        raise BundleError(f"unknown block compression {tag}")
    nrestart = struct.unpack_from("<I", body, len(body) - 4)[0]
    end = len(body) - 4 - 4 * nrestart
    p, key = 0, b""
    while p < end:
        shared, p = _varint(body, p)
        non_shared, p = _varint(body, p)
        vlen, p = _varint(body, p)
        key = key[:shared] + body[p:p + non_shared]
        p += non_shared
        yield key, body[p:p + vlen]
        p += vlen


def _fields(b: bytes):
    p = 0
    while p < len(b):
        tag, p = _varint(b, p)
        f, wt = tag >> 3, tag & 7
        if wt == 0:
            v, p = _varint(b, p)
        elif wt == 1:
            v, p = b[p:p + 8], p + 8
        elif wt == 2:
            ln, p = _varint(b, p)
            v, p = b[p:p + ln], p + ln
        elif wt == 5:
            v, p = b[p:p + 4], p + 4
        else:
            raise BundleError(f"unsupported protobuf wire type {wt}")
        yield f, wt, v


def _shape(b: bytes) -> Tuple[int, ...]:
    dims = []
    for f, wt, v in _fields(b):
        if f == 2 and wt == 2:  # TensorShapeProto.Dim
            size = 0
            for g, gw, gv in _fields(v):
                if g == 1 and gw == 0:
                    size = gv
            dims.append(size)
    return tuple(dims)


def read_index(index_path: str, verify: bool = True) -> Dict[str, dict]:
    """name -> {dtype, shape, shard_id, offset, size, crc32c}; the header entry (key '') is returned under ''."""
    buf = open(index_path, "rb").read()
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _MAGIC:
        raise BundleError(f"{index_path}: not a tensor-bundle index (bad magic)")
    foot = buf[-48:]
    p = 0
    _, p = _varint(foot, p)  # metaindex handle
    _, p = _varint(foot, p)
    ioff, p = _varint(foot, p)
    isize, p = _varint(foot, p)
    entries: Dict[str, dict] = {}
    for _, handle in _block(buf, ioff, isize, verify):
        boff, q = _varint(handle, 0)
        bsize, q = _varint(handle, q)
        for key, val in _block(buf, boff, bsize, verify):
            name = key.decode()
            if name == "":
                hdr = {"num_shards": 1}
                for f, wt, v in _fields(val):
                    if f == 1 and wt == 0:
                        hdr["num_shards"] = v
                    elif f == 2 and wt == 0:
                        hdr["endianness"] = v
                if hdr.get("endianness", 0) != 0:
                    raise BundleError("big-endian bundles are not supported")
                entries[""] = hdr
                continue
            e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
            for f, wt, v in _fields(val):
                if f == 1 and wt == 0:
                    e["dtype"] = v
                elif f == 2 and wt == 2:
                    e["shape"] = _shape(v)
                elif f == 3 and wt == 0:
                    e["shard_id"] = v
                elif f == 4 and wt == 0:
                    e["offset"] = v
                elif f == 5 and wt == 0:
                    e["size"] = v
                elif f == 6 and wt == 5:
                    e["crc32c"] = struct.unpack("<I", v)[0]
                elif f == 7:
                    e["sliced"] = True
            entries[name] = e
    return entries


def load_variables(folder: str, verify: bool = True) -> Dict[str, np.ndarray]:
    """All variables of a SavedModel folder (or of its `variables/` sub-folder) as NumPy arrays, by checkpoint key."""
    vdir = os.path.join(folder, "variables") if os.path.isdir(os.path.join(folder, "variables")) else folder
    index = os.path.join(vdir, "variables.index")
    if not os.path.exists(index):
        raise FileNotFoundError(index)
    entries = read_index(index, verify)
    nshards = entries.pop("", {"num_shards": 1})["num_shards"]
    shards: Dict[int, bytes] = {}
    out: Dict[str, np.ndarray] = {}
    for name, e in entries.items():
        if e["sliced"]:
            raise BundleError(f"{name}: partitioned variables are not supported")
        if e["dtype"] not in _DTYPES:
            continue  # strings / variants (e.g. the object graph of TF2 checkpoints) hold no weights
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open(os.path.join(vdir, f"variables.data-{sid:05d}-of-{nshards:05d}"), "rb").read()
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(_DTYPES[e["dtype"]])
        count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if len(raw) != e["size"] or count * dt.itemsize != e["size"]:
            raise BundleError(f"{name}: {e['size']} bytes on record, shape {e['shape']} of {dt} needs {count * dt.itemsize}")
        if verify and e["crc32c"] is not None and _unmask(e["crc32c"]) != crc32c(raw):
            raise BundleError(f"{name}: tensor checksum mismatch")
        out[name] = np.frombuffer(raw, dtype=dt.newbyteorder("<")).reshape(e["shape"]).copy()
    return out
