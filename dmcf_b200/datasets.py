"""Frames in and results out, either side of the hot path: the reference's ``*.msgpack.zst`` frame
files (datasets/dataset_reader_physics.py:179-207, run_sample.py:200-204), the rollout selection of ``get_rollout``
(datasets/dataset_reader_physics.py:410-456 on top of PhysicsSimDataFlow :296-352 with window=0, no shuffling, no
augmentation) and the result writer (``write_results`` :520-526).

The reference uses ``zstandard`` + ``msgpack_numpy`` + ``tensorpack`` + ``h5py``; none of them is needed here:
zstd comes from pyarrow, arrays use msgpack-numpy's plain dict encoding ``{nd, type, kind, shape, data}`` (observed in
datasets/canyon_data/canyon.msgpack.zst), and results are written as ``.npz`` with the same group/dataset names and
attributes (HDF5 only matters to utils/draw_sim2d.py)."""
from __future__ import annotations

import glob
import os

import numpy as np

FRAME_KEYS = ("pos", "vel", "grav", "m", "viscosity")


def _zstd():
    try:
        import pyarrow
        return pyarrow.Codec("zstd"), pyarrow
    except Exception as e:  # pragma: no cover
        raise RuntimeError("reading .msgpack.zst needs pyarrow (zstd codec)") from e


def _decode_hook(d):
    d = {(k.decode() if isinstance(k, bytes) else k): v for k, v in d.items()}
    if "nd" in d and "type" in d and "data" in d:
        t = d["type"].decode() if isinstance(d["type"], bytes) else d["type"]
        a = np.frombuffer(d["data"], dtype=np.dtype(t))
        return a.reshape(d["shape"]).copy() if d["nd"] else a[0]
    return d


def _encode_default(o):
    if isinstance(o, np.ndarray):
        return {"nd": True, "type": o.dtype.str, "kind": "", "shape": list(o.shape), "data": np.ascontiguousarray(o).tobytes()}
    if isinstance(o, np.generic):
        return {"nd": False, "type": o.dtype.str, "data": o.tobytes()}
    raise TypeError(type(o))


def load_msgpack_zst(path):
    """List of frame dicts (run_sample.py:200-204)."""
    import msgpack
    _, pyarrow = _zstd()
    with open(path, "rb") as fh:
        raw = fh.read()
    buf = pyarrow.CompressedInputStream(pyarrow.BufferReader(raw), "zstd").read()
    return msgpack.unpackb(buf, raw=True, object_hook=_decode_hook, strict_map_key=False)


def save_msgpack_zst(path, frames):
    """Writes frames in the encoding the reference's files use (for fixtures and exported rollouts)."""
    import msgpack
    codec, _ = _zstd()
    packed = msgpack.packb(frames, default=_encode_default, use_bin_type=True)
    with open(path, "wb") as fh:
        fh.write(codec.compress(packed, asbytes=True))


class Dataset:
    """datasets/dataset_reader_physics.py:179-207: a directory of ``*.msgpack.zst`` files, or in-memory sequences."""

    def __init__(self, data=None, dataset_path=None):
        self.data, self.files = None, None
        if dataset_path is not None:
            self.files = sorted(glob.glob(os.path.join(dataset_path, "*.msgpack.zst")))
            if not self.files:
                raise FileNotFoundError(f"no *.msgpack.zst files under {dataset_path}")
        elif data is not None:
            self.data = data
        else:
            raise NotImplementedError("generator datasets (column / free_fall) are created by the reference's generators")

    def __len__(self):
        return len(self.data) if self.data is not None else len(self.files)

    def __getitem__(self, idx):
        return self.data[idx] if self.data is not None else load_msgpack_zst(self.files[idx])


def _samples(dataset, translate=None, scale=None):
    """PhysicsSimDataFlow.__iter__ with window=0 (one frame per sample), pre_frames=0, stride=1, no shuffle / augment
    (:296-352): arrays get a leading time axis of length 1; box / box_normals come from frame 0 of the sequence."""
    for seq_i in range(len(dataset)):
        data = dataset[seq_i]
        for fi in range(len(data)):
            frame = data[fi]
            s = {}
            for k in FRAME_KEYS:
                s[k] = np.stack([np.asarray(frame[k], np.float32)], 0) if k in frame else [None]
            for k in ("box", "box_normals"):
                a = np.asarray(data[0][k], np.float32) if k in data[0] else np.empty((0, 3), np.float32)
                s[k] = a.reshape(1, -1, 3)
            for k in ("frame_id", "scene_id"):
                s[k] = np.stack([frame.get(k, None)], 0)
            if s["grav"][0] is not None:
                s["grav"] = np.full_like(s["vel"], np.expand_dims(s["grav"], 1))
            if translate is not None:
                s["pos"] = s["pos"] + np.asarray(translate, np.float32)
                s["box"] = s["box"] + np.asarray(translate, np.float32)
            if scale is not None:
                for k in ("pos", "box", "vel"):
                    s[k] = s[k] * np.float32(scale)
                if s["grav"][0] is not None:
                    s["grav"] = s["grav"] * np.float32(scale)
            yield s


def get_rollout(dataset, stride=1, time_start=0, time_end=None, random_start=1, cnt=None, rng=None, **kwargs):
    """datasets/dataset_reader_physics.py:410-456: one dict per sequence with 'pos','vel','grav','m','viscosity',
    'frame_id','scene_id','box','box_normals' concatenated over the selected frames (first axis = time)."""
    rng = rng or np.random
    rollout, random_off = [], 0
    for data in _samples(dataset, kwargs.get("translate"), kwargs.get("scale")):
        fid = int(np.asarray(data["frame_id"][0]))
        if fid == 0:
            if cnt is not None and len(rollout) >= cnt:
                break
            rollout.append([])
            random_off = rng.randint(random_start * stride) if random_start > 1 else 0
        if not rollout:
            rollout.append([])
        if fid < time_start * stride + random_off or fid % stride != 0 or (
                time_end is not None and fid >= time_end * stride + random_off):
            continue
        rollout[-1].append(data)
    out = []
    for frames in rollout:
        merge = {}
        for k in ("pos", "vel", "grav", "m", "viscosity", "frame_id", "scene_id", "box", "box_normals"):
            parts = [d[k] for d in frames]
            if parts and all(p is not None and not (isinstance(p, list) and p[0] is None) for p in parts):
                merge[k] = np.concatenate(parts, 0)
        if frames:
            out.append(merge)
    return out


def write_results(path, name, data):
    """write_results (:520-526) with ``.npz`` as the container: arrays ``<name>/<dataset>``, attributes
    ``<name>/<dataset>@type`` and ``@dim``."""
    out = {}
    for d, props in data:
        key = f"{name}/{props['name']}"
        out[key] = np.asarray(d)
        out[key + "@type"] = np.asarray(props.get("type", "DENSITY"))
        out[key + "@dim"] = np.asarray(np.asarray(d).shape)
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    np.savez_compressed(path, **out)
    return path
