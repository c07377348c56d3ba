"""Host-side orchestration of run_test / run_valid (dmcf_b200/pipeline.py, the reference's pipelines/simulator.py:110-285) with
a stub simulator: frame selection, rollout bookkeeping, result files, metric aggregation.  The GPU pieces it strings together
(Simulator.run_rollout, metrics.rollout_metrics) have their own GPU tests."""
import os

import numpy as np
import pytest
import torch

from dmcf_b200 import datasets, metrics, pipeline


class StubModel:
    name = "Stub"
    particle_radii = [0.1]
    window_dens = None

    def __call__(self, inputs, training=False):
        pos, vel = inputs[0], inputs[1]
        return pos + 0.5 * vel, vel


class StubSim:
    device = torch.device("cpu")

    def __init__(self):
        self.model = StubModel()

    def run_rollout(self, inputs, timesteps=2):  # the bookkeeping of Simulator.run_rollout without the CUDA calls
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
        samples = [[t(d["pos"][0]), t(d["vel"][0]), None, None, t(d["box"][0]), t(d["box_normals"][0])] for d in inputs]
        results = [[s] for s in samples]
        for _ in range(timesteps - 1):
            for i, s in enumerate(samples):
                pos, vel = self.model(s)
                samples[i] = [pos, vel] + s[2:]
                results[i].append(samples[i])
        return results


def make_dataset(n_seq=2, n_frames=6, n=40, seed=0):
    rng = np.random.default_rng(seed)
    seqs = []
    for s in range(n_seq):
        pos0 = rng.random((n, 3)).astype(np.float32)
        vel = (rng.standard_normal((n, 3)) * 0.01).astype(np.float32)
        box = rng.random((15, 3)).astype(np.float32) * 2 - 0.5
        frames = []
        for f in range(n_frames):
            fr = {"pos": pos0 + f * 0.5 * vel, "vel": vel, "frame_id": f, "scene_id": "s%d" % s}
            if f == 0:
                fr["box"], fr["box_normals"] = box, np.zeros_like(box)
            frames.append(fr)
        seqs.append(frames)
    return datasets.Dataset(data=seqs)


CFG = {"data_generator": {"scale": [1.0, 1.0, 1.0], "train": {"stride": 1, "repeat": True, "num_workers": 2},
                          "valid": {"stride": 1, "time_end": 5, "eval_stride": 2}, "test": {"stride": 1, "time_start": 0, "time_end": 4}},
       "output_dir": "./output"}


def test_run_test_writes_one_file_per_sequence(tmp_path):
    sim, ds = StubSim(), make_dataset()
    stale = tmp_path / "visual" / "0000"
    stale.mkdir(parents=True)
    (stale / "0003.npz").write_bytes(b"old epoch")
    written, valid = pipeline.run_test(sim, ds, CFG, str(tmp_path), epoch=7)
    assert valid is None and len(written) == 2
    assert sorted(os.listdir(stale)) == ["0007.npz"]  # only the newest epoch is kept (pipelines/simulator.py:151-155)
    z = np.load(written[1])
    assert z["Stub/pred"].shape == (4, 40, 3) and z["Stub/gt"].shape == (4, 40, 3) and z["Stub/bnd"].shape == (15, 3)
    assert str(z["Stub/pred@type"]) == "PARTICLE"
    # the stub integrates exactly what the synthetic ground truth does
    assert np.allclose(z["Stub/pred"], z["Stub/gt"], atol=1e-6)


def test_run_valid_aggregates_like_the_reference():
    sim, ds = StubSim(), make_dataset(n_seq=3, n_frames=7)
    calls = []

    def metric_fn(pos, vel, target_pos, target_vel, box, model=None, split="valid"):
        calls.append(split)
        d = float(np.mean(metrics.distance(target_pos, pos)))
        return {"mse_val": d + 1.0, "chamfer_val": 2.0}

    out = pipeline.run_valid(sim, ds, CFG, epoch=3, metric_fn=metric_fn)
    # time_end 5 -> frames 0..4, eval_stride 2 -> steps 2 and 4 of each of the 3 sequences
    assert len(calls) == 6 and set(calls) == {"valid"}
    assert abs(out["mse_val"] - 1.0) < 1e-5 and abs(out["chamfer_val"] - 2.0) < 1e-6
    assert abs(out["mse_single_val"]) < 1e-6  # one stub step from the true previous frame lands on the true frame
    assert abs(out["loss"] - (out["mse_val"] + out["chamfer_val"] + out["mse_single_val"])) < 1e-6


def test_run_test_with_metrics_and_errors(tmp_path):
    sim, ds = StubSim(), make_dataset(n_seq=1)
    cfg = dict(CFG, test_compute_metric=True)
    import dmcf_b200.metrics as m
    orig = m.rollout_metrics
    m.rollout_metrics = lambda *a, **k: {"mse_val": 0.25}
    try:
        written, valid = pipeline.run_test(sim, ds, cfg, str(tmp_path), epoch=1)
    finally:
        m.rollout_metrics = orig
    assert len(written) == 1 and abs(valid["mse_val"] - 0.25) < 1e-9 and "loss" in valid
    with pytest.raises(NotImplementedError):
        pipeline.open_split({"name": "CConvData3D"}, "test")
    with pytest.raises(FileNotFoundError):
        pipeline.open_split({"dataset_path": str(tmp_path / "nowhere")}, "test")


def test_open_split_reads_a_directory_of_frame_files(tmp_path):
    frames = make_dataset(n_seq=1).data[0]
    os.makedirs(tmp_path / "test")
    datasets.save_msgpack_zst(str(tmp_path / "test" / "sim_0001.msgpack.zst"), frames)
    ds = pipeline.open_split({"dataset_path": str(tmp_path)}, "test")
    assert len(ds) == 1
    ro = datasets.get_rollout(ds, time_end=3)
    assert ro[0]["pos"].shape == (3, 40, 3) and ro[0]["box"].shape[1:] == (15, 3)
