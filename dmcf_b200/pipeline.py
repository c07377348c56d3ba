"""The inference-side flows of the reference's pipeline around the hot path: ``run_test`` (pipelines/simulator.py:110-158:
roll out the test split, write <out_dir>/visual/<seq>/<epoch> with pred / gt / bnd) and ``run_valid`` (:160-285: roll out
the validation split, per-frame metric dicts, per-sequence and overall means, ``loss`` = sum of the means).  Everything here is
host-side orchestration of pieces that are tested on their own: ``datasets.get_rollout``, ``Simulator.run_rollout``,
``metrics.rollout_metrics``, ``datasets.write_results``.  ``run_train`` is out of scope."""
from __future__ import annotations

import glob
import logging
import os

import numpy as np
import torch

from . import datasets, metrics

log = logging.getLogger(__name__)


def _np(a):
    return a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)


def open_split(dataset_cfg, split):
    """DatasetGroup for a dataset on disk (datasets/dataset_reader_physics.py:117-140): <dataset_path>/<split>/*.msgpack.zst."""
    path = dataset_cfg.get("dataset_path")
    if not path:
        raise NotImplementedError("generator datasets (type: column / free_fall) are produced by the reference's generators; "
                                  "point dataset.dataset_path at their output")
    sub = os.path.join(path, split)
    return datasets.Dataset(dataset_path=sub if os.path.isdir(sub) else path)


def _generator_args(pipeline_cfg, split):
    gen = dict(pipeline_cfg.get("data_generator") or {})
    per_split = dict(gen.pop(split, None) or {})
    for other in ("train", "valid", "test"):
        gen.pop(other, None)
    gen.update(per_split)
    # options the reference forwards to PhysicsSimDataFlow that CHANGE the sequences (sub-sampled particles, dropped leading
    # frames, rotated gravity): not implemented here -- refuse instead of silently rolling out different data
    for k, neutral in (("sample_cnt", (None, 0)), ("pre_frames", (None, 0)), ("grav_eqvar", (None, False))):
        if gen.get(k) not in neutral:
            raise NotImplementedError(f"data_generator option {k}={gen[k]!r} is not supported by get_rollout")
    for k in ("repeat", "shuffle_buffer", "is2d", "num_workers", "sample_cnt", "augment", "eval_stride", "batch_size", "window",
              "pre_frames", "grav_eqvar"):
        gen.pop(k, None)  # data-flow options of the training loader, not of get_rollout
    if isinstance(gen.get("scale"), (list, tuple)):
        gen["scale"] = np.asarray(gen["scale"], np.float32)
    return gen, per_split


def run_test(sim, dataset, pipeline_cfg, out_dir, epoch=0, compute_metric=None, valid_dataset=None):
    """pipelines/simulator.py:110-158.  Returns the list of written files (and the validation dict when
    ``test_compute_metric`` is set, :157-158).  The reference's run_valid always rolls out ``dataset.valid``: pass that split
    as ``valid_dataset``; without it the metrics are computed on the test split (logged as a deviation)."""
    gen, _ = _generator_args(pipeline_cfg, "test")
    test_data = datasets.get_rollout(dataset, **gen)
    if not test_data:
        raise ValueError("the test split holds no sequence")
    results = sim.run_rollout(test_data, test_data[0]["pos"].shape[0])
    written = []
    for i, data in enumerate(test_data):
        pos = np.stack([_np(r[0]) for r in results[i]])
        seq_dir = os.path.join(out_dir, "visual", "%04d" % i)
        os.makedirs(seq_dir, exist_ok=True)
        target = os.path.join(seq_dir, "%04d.npz" % epoch)
        written.append(datasets.write_results(target, sim.model.name,
                                              [(pos, {"name": "pred", "type": "PARTICLE"}),
                                               (np.asarray(data["pos"]), {"name": "gt", "type": "PARTICLE"}),
                                               (np.asarray(data["box"][0]), {"name": "bnd", "type": "PARTICLE"})]))
        for f in glob.glob(os.path.join(seq_dir, "*.npz")):  # :151-155 keeps only the newest epoch
            if os.path.abspath(f) != os.path.abspath(target):
                os.remove(f)
    valid = None
    if compute_metric if compute_metric is not None else pipeline_cfg.get("test_compute_metric", False):
        if valid_dataset is None:
            log.warning("test_compute_metric: no validation split given, computing the metrics on the test split")
        valid = run_valid(sim, valid_dataset if valid_dataset is not None else dataset, pipeline_cfg, epoch,
                          split="valid" if valid_dataset is not None else "test")
    return written, valid


def run_valid(sim, dataset, pipeline_cfg, epoch=0, split="valid", metric_fn=None):
    """pipelines/simulator.py:160-285.  Returns the dict of metric means over all evaluated frames plus ``loss`` = their sum."""
    metric_fn = metric_fn or metrics.rollout_metrics
    gen, per_split = _generator_args(pipeline_cfg, "valid")
    eval_stride = int(per_split.get("eval_stride", 1))
    valid_data = datasets.get_rollout(dataset, **gen)
    if not valid_data:
        raise ValueError("the validation split holds no sequence")
    results = sim.run_rollout(valid_data, valid_data[0]["pos"].shape[0])
    dev = getattr(sim, "device", None)
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    losses = []
    for i, data in enumerate(valid_data):
        target_pos, target_vel = data["pos"], data["vel"]
        box = np.asarray(data["box"][0], np.float32)
        loss_seq = []
        for step in range(1, min(target_pos.shape[0], len(results[i]))):
            if step % eval_stride != 0:
                continue
            pos, vel = results[i][step][:2]
            loss = dict(metric_fn(pos, vel, target_pos[step], target_vel[step], box, model=sim.model, split=split))
            # single-step error from the ground-truth previous frame (:251-256)
            with torch.no_grad():
                pos_sub = sim.model([t(target_pos[step - 1]), t(target_vel[step - 1])] + list(results[i][step][2:]))[0]
            loss["mse_single_val"] = float(np.mean(metrics.distance(target_pos[step], pos_sub)))
            losses.append(loss)
            loss_seq.append(loss)
        if loss_seq:
            mean = metrics.merge_dicts(loss_seq, lambda x, y: x + y / len(loss_seq))
            log.info("%d - %s", i, " ".join("%s: %.05f" % kv for kv in mean.items()))
    if not losses:
        raise ValueError("no frame was evaluated (sequences shorter than two frames, or eval_stride too large)")
    out = metrics.merge_dicts(losses, lambda x, y: x + y / len(losses))
    out["loss"] = float(sum(out.values()))
    log.info("validation of epoch %d - %s", epoch, " ".join("%s: %.05f" % kv for kv in out.items()))
    return out
