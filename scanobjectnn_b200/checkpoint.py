"""Checkpoints for the VariableStore: the reference restores its weights with ``tf.train.Saver().restore(sess, MODEL_PATH)``
(pointnet2/evaluate_scenennobjects.py:131-141) -- a TensorFlow V2 "tensor bundle" (``model.ckpt.index`` +
``model.ckpt.data-00000-of-00001``).  The store's keys ARE the TF variable names (``layer1/conv0/weights``,
``fc1/bn/moving_mean`` ...), so loading is a name-for-name copy.

* ``read_tf_checkpoint(prefix)``  pure-Python reader of the bundle (no TensorFlow needed, none is installed here):
  the ``.index`` file is an SSTable (leveldb table format: prefix-compressed key/value blocks, block handles, 48-byte
  footer with magic 0xdb4775248b80fb57) whose values are ``BundleEntryProto`` messages (dtype, shape, shard, offset, size);
  tensor bytes sit raw in the data shards.  Restated from the published format (tensorflow/core/util/tensor_bundle,
  tensorflow/core/lib/io/{table,block,format}.cc); checked here against ``write_tf_checkpoint`` (same spec, so the pair is
  self-consistent) -- NOT against a file written by TensorFlow itself (unavailable offline).
* ``write_tf_checkpoint(prefix, tensors)``  the matching writer (export back to the reference).
* ``load_npz`` / ``save_npz``  plain numpy archives with the same names.
* ``load_into(params, tensors)``  copy into a VariableStore (shape-checked; BN / optimizer slots the model does not have are
  reported, not silently dropped) and invalidate its folded-weight caches.
"""
from __future__ import annotations

import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}
_DTYPE_CODE = {np.dtype(v): k for k, v in _DTYPES.items()}


# ---------------------------------------------------------------------------------------------------------
# varints / protobuf wire format (only what BundleHeaderProto / BundleEntryProto need)
# ---------------------------------------------------------------------------------------------------------
def _get_varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_fields(buf):
    """protobuf message -> list of (field number, wire type, value)"""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _get_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]; pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            v = buf[pos:pos + n]; pos += n
        elif wt == 5:
            v = buf[pos:pos + 4]; pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        out.append((fno, wt, v))
    return out


def _parse_entry(buf):
    """BundleEntryProto: 1 dtype, 2 shape {2: dim {1: size}}, 3 shard_id, 4 offset, 5 size, 6 crc32c (fixed32)"""
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0)
    for fno, wt, v in _parse_fields(buf):
        if fno == 1:
            e["dtype"] = v
        elif fno == 2:
            for f2, _, dimbuf in _parse_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, sv in _parse_fields(dimbuf):
                        if f3 == 1:
                            size = sv
                    e["shape"].append(size)
        elif fno == 3:
            e["shard_id"] = v
        elif fno == 4:
            e["offset"] = v
        elif fno == 5:
            e["size"] = v
    return e


def _field(fno, wt, payload: bytes) -> bytes:
    return _put_varint((fno << 3) | wt) + payload


def _entry_bytes(dtype_code, shape, offset, size, crc) -> bytes:
    dims = b"".join(_field(2, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(int(s))))) for s in shape)
    out = _field(1, 0, _put_varint(dtype_code)) + _field(2, 2, _put_varint(len(dims)) + dims)
    if offset:
        out += _field(4, 0, _put_varint(offset))
    out += _field(5, 0, _put_varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return out


# ---------------------------------------------------------------------------------------------------------
# crc32c (Castagnoli), masked as leveldb / TensorFlow store it
# ---------------------------------------------------------------------------------------------------------
_CRC_TABLE = None


def _crc32c(data: bytes) -> int:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl.append(c)
        _CRC_TABLE = np.array(tbl, dtype=np.uint32)
    crc = 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in data:
        crc = int(tbl[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------------------
# SSTable blocks
# ---------------------------------------------------------------------------------------------------------
def _read_block(buf, offset, size):
    """entries of one block -> list of (key bytes, value bytes)"""
    blk = buf[offset:offset + size]
    n_restarts = struct.unpack("<I", blk[-4:])[0]
    end = len(blk) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _get_varint(blk, pos)
        non_shared, pos = _get_varint(blk, pos)
        vlen, pos = _get_varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        out.append((key, blk[pos:pos + vlen]))
        pos += vlen
    return out


def _build_block(items, restart_interval=16) -> bytes:
    out, restarts, last = bytearray(), [], b""
    for i, (k, v) in enumerate(items):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
        last = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _handle(offset, size) -> bytes:
    return _put_varint(offset) + _put_varint(size)


def read_tf_checkpoint(prefix: str) -> dict:
    """``prefix`` as passed to tf.train.Saver.restore (e.g. ".../model.ckpt") -> {variable name: numpy array}."""
    with open(prefix + ".index", "rb") as f:
        idx = f.read()
    if len(idx) < 48 or struct.unpack("<Q", idx[-8:])[0] != _MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorFlow V2 checkpoint index (bad table magic)")
    footer = idx[-48:]
    _, p = _get_varint(footer, 0)          # metaindex handle (offset, size): unused
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isz, p = _get_varint(footer, p)
    entries = {}
    num_shards = 1
    for _, hv in _read_block(idx, ioff, isz):
        boff, q = _get_varint(hv, 0)
        bsz, _ = _get_varint(hv, q)
        if idx[boff + bsz] != 0:
            raise ValueError("compressed index blocks are not supported (TensorFlow writes them uncompressed)")
        for k, v in _read_block(idx, boff, bsz):
            if k == b"":
                for fno, _, hvv in _parse_fields(v):      # BundleHeaderProto: 1 num_shards, 2 endianness, 3 version
                    if fno == 1:
                        num_shards = hvv
                    if fno == 2 and hvv != 0:
                        raise ValueError("big-endian checkpoints are not supported")
            else:
                entries[k.decode("utf-8")] = _parse_entry(v)
    shards = {}
    out = {}
    for name, e in entries.items():
        if e["dtype"] not in _DTYPES:
            continue                                       # strings etc. (e.g. the saver's metadata): not variables
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap(f"{prefix}.data-{sid:05d}-of-{num_shards:05d}", dtype=np.uint8, mode="r")
        raw = shards[sid][e["offset"]:e["offset"] + e["size"]]
        out[name] = np.frombuffer(raw.tobytes(), dtype=_DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    return out


def write_tf_checkpoint(prefix: str, tensors: dict) -> None:
    """Write {name: array} as a single-shard V2 bundle readable by ``read_tf_checkpoint`` (and, per the format, by
    tf.train.Saver / tf.train.load_checkpoint)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    items = []
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name])
            a = a.copy(order="C") if not a.flags.c_contiguous else a      # (ascontiguousarray would turn scalars into 1-D)
            if a.dtype not in _DTYPE_CODE:
                raise TypeError(f"{name}: dtype {a.dtype} is not supported")
            raw = a.tobytes()
            f.write(raw)
            items.append((name.encode("utf-8"), _entry_bytes(_DTYPE_CODE[a.dtype], a.shape, offset, len(raw), _mask(_crc32c(raw)))))
            offset += len(raw)
    header = _field(1, 0, _put_varint(1)) + _field(3, 2, (lambda d: _put_varint(len(d)) + d)(_field(1, 0, _put_varint(1))))
    items = [(b"", header)] + items
    out = bytearray()

    def emit(block: bytes):
        off = len(out)
        out.extend(block)
        out.extend(b"\x00" + struct.pack("<I", _mask(_crc32c(block + b"\x00"))))
        return off, len(block)

    doff, dsz = emit(_build_block(items))
    moff, msz = emit(_build_block([]))
    ioff, isz = emit(_build_block([(items[-1][0] + b"\x00", _handle(doff, dsz))], restart_interval=1))
    footer = _handle(moff, msz) + _handle(ioff, isz)
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC))
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


# ---------------------------------------------------------------------------------------------------------
# VariableStore <-> files
# ---------------------------------------------------------------------------------------------------------
def load_npz(path: str) -> dict:
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def save_npz(params, path: str) -> None:
    np.savez(path, **{k: v.detach().cpu().numpy() for k, v in params.items()})


def load_into(params, tensors: dict, strict: bool = True):
    """Copy {TF name: array} into the store.  Names the model has but the checkpoint lacks -> KeyError (strict) or kept;
    extra checkpoint names (optimizer slots ``.../Adam``, ``beta1_power``, the ``batch`` counter) are returned, not loaded.
    In-place copies: views held by a trainer's flat bucket or by engines see the new values; folded caches are dropped."""
    import torch
    missing = [k for k in params.keys() if k not in tensors]
    if missing and strict:
        raise KeyError(f"checkpoint lacks {len(missing)} variables, e.g. {missing[:4]}")
    unused = []
    for name, arr in tensors.items():
        if name not in params:
            unused.append(name)
            continue
        dst = params[name]
        src = torch.as_tensor(np.asarray(arr))
        if tuple(src.shape) != tuple(dst.shape):
            if src.numel() == dst.numel():
                src = src.reshape(dst.shape)             # e.g. conv kernels stored (1,1,Cin,Cout) vs (Cin,Cout)
            else:
                raise ValueError(f"{name}: checkpoint shape {tuple(src.shape)} != model shape {tuple(dst.shape)}")
        dst.copy_(src.to(dst.dtype))
    params.invalidate()
    return unused


def restore(params, source: str, strict: bool = True):
    """``source``: an ``.npz`` archive or a TF checkpoint prefix (``.../model.ckpt``)."""
    tensors = load_npz(source) if source.endswith(".npz") else read_tf_checkpoint(source)
    return load_into(params, tensors, strict=strict)
