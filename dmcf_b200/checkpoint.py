"""TensorFlow-free reader for the TF2 object-graph checkpoints DMCF ships (checkpoints/*/ckpt.index +
ckpt.data-00000-of-00001), the format written by ``tf.train.Checkpoint`` in pipelines/base_pipeline.py:155-191.

``ckpt.index`` is a LevelDB-format table (uncompressed blocks, prefix-compressed keys) whose values are
``BundleEntryProto`` messages {1: dtype, 2: shape, 3: shard_id, 4: offset, 5: size, 6: crc32c}; the data file holds
raw little-endian tensors at those offsets.
"""
from __future__ import annotations

import os
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}


def _varint(buf, pos):
    result, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _block_entries(block):
    """Yields (key, value) of one table block (restart array at the tail is not needed for a linear scan)."""
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _read_block(data, offset, size):
    if data[offset + size] != 0:
        raise ValueError("compressed checkpoint index blocks are not supported")
    return memoryview(data)[offset:offset + size]


def _parse_proto(buf):
    """Minimal protobuf wire parser -> {field: [values]} (varints as int, length-delimited as bytes)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
        elif wire == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wire}")
        out.setdefault(field, []).append(v)
    return out


def _shape(buf):
    dims = []
    for d in _parse_proto(buf).get(2, []):
        dims.append(_parse_proto(d).get(1, [0])[0])
    return tuple(int(x) for x in dims)


def read_index(prefix):
    """{key: (dtype_enum, shape, offset, size)} for every tensor of the bundle at ``prefix`` (e.g. '.../ckpt')."""
    with open(prefix + ".index", "rb") as fh:
        data = fh.read()
    if struct.unpack_from("<Q", data, len(data) - 8)[0] != _MAGIC:
        raise ValueError(f"{prefix}.index is not a TensorFlow bundle index")
    footer = data[-48:]
    _, p = _varint(footer, 0)      # metaindex handle offset
    _, p = _varint(footer, p)      # metaindex handle size
    idx_off, p = _varint(footer, p)
    idx_size, p = _varint(footer, p)
    entries = {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size)):
        off, q = _varint(handle, 0)
        size, q = _varint(handle, q)
        for key, value in _block_entries(_read_block(data, off, size)):
            if key == b"":
                continue  # BundleHeaderProto
            msg = _parse_proto(value)
            entries[key.decode()] = (msg.get(1, [0])[0], _shape(msg[2][0]) if 2 in msg else (), msg.get(4, [0])[0],
                                     msg.get(5, [0])[0], msg.get(3, [0])[0])
    return entries


def resolve_prefix(path):
    """Bundle prefix for ``path``: the prefix itself, or for a directory the bundle named by the CheckpointManager state file
    ``checkpoint`` (``model_checkpoint_path: "ckpt-12"``, what ``manager.latest_checkpoint`` restores,
    pipelines/base_pipeline.py:155-187), else the ``*.index`` with the largest trailing integer (ckpt-10 after ckpt-9)."""
    import re
    if not os.path.isdir(path):
        return path
    state = os.path.join(path, "checkpoint")
    if os.path.exists(state):
        with open(state) as fh:
            m = re.search(r'^model_checkpoint_path:\s*"([^"]+)"', fh.read(), re.M)
        if m:
            cand = m.group(1) if os.path.isabs(m.group(1)) else os.path.join(path, m.group(1))
            if os.path.exists(cand + ".index"):
                return cand
    cands = [f[:-6] for f in os.listdir(path) if f.endswith(".index")]
    if not cands:
        raise FileNotFoundError(f"no *.index in {path}")

    def order(name):
        nums = re.findall(r"\d+", name)
        return (int(nums[-1]) if nums else -1, name)
    return os.path.join(path, max(cands, key=order))


def load_checkpoint(path, return_prefix=False):
    """Returns {variable path without '/.ATTRIBUTES/VARIABLE_VALUE': ndarray}; optimizer slots are dropped.
    ``path`` may be the bundle prefix ('.../ckpt') or the directory holding 'ckpt.index' (run_sample.py:184-197)."""
    prefix = resolve_prefix(path)
    entries = read_index(prefix)
    with open(prefix + ".data-00000-of-00001", "rb") as fh:
        blob = fh.read()
    out = {}
    suffix = "/.ATTRIBUTES/VARIABLE_VALUE"
    for key, (dt, shape, off, size, _shard) in entries.items():
        if not key.endswith(suffix) or ".OPTIMIZER_SLOT" in key or dt not in _DTYPES:
            continue
        if entries[key][4] != 0:
            raise NotImplementedError(f"{key}: tensor lives in data shard {entries[key][4]}; only single-shard bundles are read")
        arr = np.frombuffer(blob, dtype=np.dtype(_DTYPES[dt]).newbyteorder("<"), count=int(np.prod(shape, dtype=np.int64)),
                            offset=off).reshape(shape)
        out[key[:-len(suffix)]] = arr.copy()
    return (out, prefix) if return_prefix else out


def model_weights(ckpt):
    """Strips the leading 'model/' of the object-graph paths and drops non-model entries."""
    return {k[len("model/"):]: v for k, v in ckpt.items() if k.startswith("model/")}
