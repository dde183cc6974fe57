"""TF checkpoint import of the DM21 weights (grad_dft/functional.py:824-928) without TensorFlow: the tensor-bundle
reader against the reference's own checkpoint index (committed, 681 bytes) and -- where the reference tree is present
-- against the checkpoint data (CRC-32C of every tensor as TensorFlow wrote it) and an independent NumPy forward."""
import os
from pathlib import Path

import numpy as np
import pytest
import torch

from graddft_b200 import tf_bundle
from graddft_b200.functional import DM21, dm21_checkpoint_params

GOLD = Path(__file__).parent / "golden"
REF = "/root/reference/models/DM21_model"


def test_crc32c_known_answers():
    # RFC 3720 B.4 test vectors
    assert tf_bundle.crc32c(b"") == 0
    assert tf_bundle.crc32c(bytes(32)) == 0x8A9136AA
    assert tf_bundle.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert tf_bundle.crc32c(bytes(range(32))) == 0x46DD794E
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283


def test_snappy_roundtrip_elements():
    # literal "abcd", copy-1 (offset 4, len 4), literal "xy", copy-2 (offset 2, len 6: overlapping run)
    stream = bytes([16]) + bytes([3 << 2]) + b"abcd" + bytes([(0 << 5) | (0 << 2) | 1, 4]) + bytes([1 << 2]) + b"xy" + bytes([(5 << 2) | 2, 2, 0])
    assert tf_bundle.snappy_uncompress(stream) == b"abcdabcdxyxyxyxy"
    with pytest.raises(tf_bundle.BundleError):
        tf_bundle.snappy_uncompress(bytes([4]) + bytes([(0 << 2) | 1, 9]))


def test_index_of_reference_checkpoint(tmp_path):
    """Block checksums verify, names/shapes/offsets are the DM21 architecture of functional.py:793-822."""
    idx = tf_bundle.read_index(str(GOLD / "dm21_variables.index"), verify=True)
    assert idx.pop("")["num_shards"] == 1
    assert len(idx) == 28
    base = "hub_wrapper/local_functional_v2/"
    assert idx[base + "SquashUnprocessedData/linear/w"]["shape"] == (11, 256)
    assert idx[base + "OutputLayer/linear/w"]["shape"] == (256, 3)
    for k in ["", "_1", "_2", "_3", "_4", "_5"]:
        assert idx[base + f"MLP/ResidualBlock{k}/linear/w"]["shape"] == (256, 256)
        assert idx[base + f"MLP/ResidualBlock{k}/layer_norm/gamma"]["shape"] == (256,)
    assert all(e["dtype"] == 1 for e in idx.values())  # DT_FLOAT
    spans = sorted((e["offset"], e["size"]) for e in idx.values())
    assert spans[0][0] == 0 and all(a + s == b for (a, s), (b, _) in zip(spans, spans[1:]))
    assert spans[-1][0] + spans[-1][1] == 1606668  # size of variables.data-00000-of-00001
    bad = bytearray((GOLD / "dm21_variables.index").read_bytes())
    bad[40] ^= 1
    (tmp_path / "variables.index").write_bytes(bytes(bad))
    with pytest.raises(tf_bundle.BundleError):
        tf_bundle.read_index(str(tmp_path / "variables.index"))


def test_seeded_weights_unchanged_by_the_folder_argument():
    a = DM21().generate_DM21_weights(seed=7)
    b = DM21().generate_DM21_weights(None, seed=7)
    assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
    k1 = a["Dense_1.kernel"]
    assert abs(float(torch.diagonal(k1).mean()) - 1.0) < 0.05  # identity added to the square kernels


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
def test_dm21_weights_from_the_reference_checkpoint():
    raw = tf_bundle.load_variables(REF, verify=True)  # verifies TensorFlow's per-tensor CRC-32C
    assert sum(a.size for a in raw.values()) == 401667
    fun = DM21()
    p = fun.generate_DM21_weights(REF)
    assert len(p) == 8 * 2 + 6 * 2
    base = "hub_wrapper/local_functional_v2/"
    assert torch.equal(p["Dense_0.kernel"], torch.from_numpy(raw[base + "SquashUnprocessedData/linear/w"].astype(np.float64)))
    assert torch.equal(p["Dense_3.kernel"], torch.from_numpy(raw[base + "MLP/ResidualBlock_2/linear/w"].astype(np.float64)))
    assert torch.equal(p["LayerNorm_0.scale"], torch.from_numpy(raw[base + "MLP/ResidualBlock/layer_norm/gamma"].astype(np.float64)))
    assert torch.equal(p["Dense_7.bias"], torch.from_numpy(raw[base + "OutputLayer/linear/b"].astype(np.float64)))
    # upstream's merge rule: layers beyond num_layers_with_dm_parameters keep their initial values (+ identity)
    q = fun.generate_DM21_weights(REF, num_layers_with_dm_parameters=3, seed=5)
    init = fun.generate_DM21_weights(seed=5)
    assert torch.equal(q["Dense_3.kernel"], p["Dense_3.kernel"]) and torch.equal(q["Dense_4.kernel"], init["Dense_4.kernel"])
    assert torch.equal(q["LayerNorm_2.scale"], p["LayerNorm_2.scale"]) and torch.equal(q["LayerNorm_4.scale"], init["LayerNorm_4.scale"])
    # shape mismatch (other feature count): Dense_0 keeps its initial values, the trunk is still imported
    r = fun.generate_DM21_weights(REF, n_input_features=7, seed=5)
    assert r["Dense_0.kernel"].shape == (7, 256) and torch.equal(r["Dense_2.kernel"], p["Dense_2.kernel"])
    # the network through the host mirror (CPU tensors take the composite path) against a NumPy forward from the raw arrays
    x = torch.rand(64, 11, dtype=torch.float64, generator=torch.Generator().manual_seed(1)) * 3 - 1
    got = fun.apply(p, x).numpy()
    h = np.log(np.abs(x.numpy()) + 1e-4)
    h = np.tanh(h @ raw[base + "SquashUnprocessedData/linear/w"].astype(np.float64) + raw[base + "SquashUnprocessedData/linear/b"])
    for k in ["", "_1", "_2", "_3", "_4", "_5"]:
        blk = base + f"MLP/ResidualBlock{k}/"
        y = h @ raw[blk + "linear/w"].astype(np.float64) + raw[blk + "linear/b"] + h
        mu = y.mean(-1, keepdims=True)
        y = (y - mu) / np.sqrt(((y - mu) ** 2).mean(-1, keepdims=True) + 1e-6) * raw[blk + "layer_norm/gamma"] + raw[blk + "layer_norm/beta"]
        h = np.where(y > 0, y, np.expm1(y))
    o = h @ raw[base + "OutputLayer/linear/w"].astype(np.float64) + raw[base + "OutputLayer/linear/b"]
    want = 2.0 / (1.0 + np.exp(-o / 2.0))
    assert np.abs(got - want).max() < 1e-12
    assert dm21_checkpoint_params(REF).keys() == p.keys()
