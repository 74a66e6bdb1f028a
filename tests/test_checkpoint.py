"""Checkpoint I/O (CPU): the pure-Python TensorFlow-bundle reader against the matching writer, prefix-compressed keys over
several restart intervals, npz round trip, and restoring into a VariableStore (pointnet2/evaluate_scenennobjects.py:131-141)."""
import struct

import numpy as np
import pytest

from scanobjectnn_b200 import checkpoint as ck
from scanobjectnn_b200 import pointnet2_cls_ssg


def test_crc32c_known_answers():
    # RFC 3720 test vectors for CRC-32C
    assert ck._crc32c(b"\x00" * 32) == 0x8A9136AA
    assert ck._crc32c(b"\xff" * 32) == 0x62A8AB43
    assert ck._crc32c(bytes(range(32))) == 0x46DD794E
    assert ck._crc32c(b"123456789") == 0xE3069283


def test_tf_bundle_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {f"layer{i}/conv{j}/weights": rng.standard_normal((1, 1, 3 + i, 8 + j)).astype(np.float32) for i in range(6) for j in range(5)}
    tensors["layer1/conv0/bn/moving_mean"] = rng.standard_normal(8).astype(np.float32)
    tensors["batch"] = np.array(1234.0, dtype=np.float32)          # scalar (the reference's global step variable)
    tensors["counts"] = np.arange(7, dtype=np.int64)
    prefix = str(tmp_path / "model.ckpt")
    ck.write_tf_checkpoint(prefix, tensors)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57    # leveldb table magic
    got = ck.read_tf_checkpoint(prefix)
    assert set(got) == set(tensors)
    for k, v in tensors.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and np.array_equal(got[k], v), k


def test_restore_into_variable_store(tmp_path):
    p = pointnet2_cls_ssg.init_params(seed=3, device="cpu", randomize_bn=True)
    src = {k: (v.numpy() * 1.5 + 0.25).astype(np.float32) for k, v in p.items()}
    src["fc1/weights/Adam"] = np.zeros((1024, 512), np.float32)      # optimizer slot: reported, not loaded
    prefix = str(tmp_path / "best_model.ckpt")
    ck.write_tf_checkpoint(prefix, src)
    p.folded("fc1")                                                   # populate the folded cache: must be dropped by the load
    assert p._cache
    unused = ck.restore(p, prefix)
    assert unused == ["fc1/weights/Adam"] and not p._cache
    for k in p:
        assert np.array_equal(p[k].numpy(), src[k]), k
    ck.save_npz(p, str(tmp_path / "w.npz"))
    q = pointnet2_cls_ssg.init_params(seed=9, device="cpu")
    ck.restore(q, str(tmp_path / "w.npz"))
    assert all(np.array_equal(q[k].numpy(), src[k]) for k in q)
    del src["layer2/conv1/weights"]
    with pytest.raises(KeyError):
        ck.load_into(q, src)


def test_assignment_invalidates_folded_cache():
    import torch
    p = pointnet2_cls_ssg.init_params(seed=1, device="cpu")
    p.folded("fc1")
    assert p._cache
    p["fc1/weights"] = torch.zeros_like(p["fc1/weights"])
    assert not p._cache
