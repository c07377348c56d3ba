"""Frame files and rollout selection (dmcf_b200/datasets.py vs datasets/dataset_reader_physics.py:179-207, 410-456)."""
import os

import numpy as np

from dmcf_b200 import datasets


def _frames(n_frames, n, seed, scene_id="s0"):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n_frames):
        f = {"pos": rng.random((n, 3)).astype(np.float32), "vel": rng.random((n, 3)).astype(np.float32),
             "m": np.full(n, 0.125, np.float32), "viscosity": np.full(n, 0.01, np.float32),
             "grav": np.array([0, -9.81, 0], np.float32), "frame_id": np.int64(i), "scene_id": scene_id}
        if i == 0:
            f["box"] = rng.random((7, 3)).astype(np.float32)
            f["box_normals"] = rng.random((7, 3)).astype(np.float32)
        out.append(f)
    return out


def test_msgpack_zst_round_trip(tmp_path):
    frames = _frames(3, 11, 0)
    p = os.path.join(tmp_path, "seq.msgpack.zst")
    datasets.save_msgpack_zst(p, frames)
    back = datasets.load_msgpack_zst(p)
    assert len(back) == 3 and set(back[0]) == set(frames[0]) and set(back[1]) == set(frames[1])
    for a, b in zip(frames, back):
        for k in ("pos", "vel", "m", "viscosity", "grav"):
            assert b[k].dtype == a[k].dtype and np.array_equal(a[k], b[k])
        assert int(b["frame_id"]) == int(a["frame_id"])
    ds = datasets.Dataset(dataset_path=str(tmp_path))
    assert len(ds) == 1 and np.array_equal(ds[0][0]["box"], frames[0]["box"])


def test_get_rollout_selection_and_layout(tmp_path):
    for i in range(2):
        datasets.save_msgpack_zst(os.path.join(tmp_path, f"seq{i}.msgpack.zst"), _frames(10, 5 + i, i, f"s{i}"))
    ds = datasets.Dataset(dataset_path=str(tmp_path))
    ro = datasets.get_rollout(ds, stride=2, time_start=1, time_end=4)
    assert len(ro) == 2
    for i, r in enumerate(ro):
        # frames with id % 2 == 0 and 2 <= id < 8
        assert r["frame_id"].tolist() == [2, 4, 6]
        assert r["pos"].shape == (3, 5 + i, 3) and r["vel"].shape == (3, 5 + i, 3) and r["m"].shape == (3, 5 + i)
        assert r["box"].shape == (3, 7, 3) and np.array_equal(r["box"][0], r["box"][2])   # frame 0's box, repeated
        assert r["grav"].shape == (3, 5 + i, 3) and np.allclose(r["grav"][..., 1], -9.81)
    assert len(datasets.get_rollout(ds, cnt=1)) == 1
    ro = datasets.get_rollout(ds, cnt=1, scale=2.0, translate=[1.0, 0.0, 0.0])
    raw = ds[0]
    assert np.allclose(ro[0]["pos"][0], (raw[0]["pos"] + np.array([1.0, 0, 0], np.float32)) * 2.0)
    assert np.allclose(ro[0]["vel"][3], raw[3]["vel"] * 2.0)


def test_reads_the_reference_encoding_fixture():
    """tests/golden/canyon_crop.npz was cut from datasets/canyon_data/canyon.msgpack.zst with the same decoder
    (scripts/make_golden.py); a file written by save_msgpack_zst from it must decode to the same arrays."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "canyon_crop.npz"))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "c.msgpack.zst")
        datasets.save_msgpack_zst(p, [{"pos": z["pos"], "vel": z["vel"], "box": z["box"], "box_normals": z["box_normals"],
                                       "frame_id": np.int64(0), "scene_id": "canyon"}])
        f = datasets.load_msgpack_zst(p)[0]
        assert np.array_equal(f["pos"], z["pos"]) and np.array_equal(f["box_normals"], z["box_normals"])
        out = datasets.write_results(os.path.join(d, "r", "0001.npz"), "SymNet",
                                     [(f["pos"][None], {"name": "pred", "type": "PARTICLE"})])
        r = np.load(out)
        assert r["SymNet/pred"].shape == (1,) + z["pos"].shape and str(r["SymNet/pred@type"]) == "PARTICLE"


def test_checkpoint_directory_resolves_like_checkpoint_manager(tmp_path):
    """pipelines/base_pipeline.py:155-187 restores manager.latest_checkpoint: the bundle named by the `checkpoint` state file,
    else the highest trailing number (ckpt-10 after ckpt-9, not lexicographic)."""
    from dmcf_b200.checkpoint import resolve_prefix
    for n in (2, 9, 10):
        (tmp_path / f"ckpt-{n}.index").write_bytes(b"")
    assert resolve_prefix(str(tmp_path)).endswith("ckpt-10")
    (tmp_path / "checkpoint").write_text('model_checkpoint_path: "ckpt-9"\nall_model_checkpoint_paths: "ckpt-2"\n')
    assert resolve_prefix(str(tmp_path)).endswith("ckpt-9")
    assert resolve_prefix(str(tmp_path / "ckpt-2")) == str(tmp_path / "ckpt-2")
