"""Rollout pipeline: the inference half of the reference's ``Simulator`` (pipelines/simulator.py:37-109) and
checkpoint restore (pipelines/base_pipeline.py:155-187, run_sample.py:184-197).  Training / validation / logging
orchestration is out of scope."""
from __future__ import annotations

import logging
import time

import numpy as np
import torch

from .checkpoint import load_checkpoint, model_weights

log = logging.getLogger(__name__)


def _to_dev(a, device):
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


class Simulator:
    """Pipeline for the trainable simulator (inference only)."""

    def __init__(self, model, dataset=None, name="Simulator", main_log_dir="./logs/", device="cuda", split="test",
                 **kwargs):
        self.model = model
        self.dataset = dataset
        self.name = name
        self.device = torch.device(device if device != "gpu" else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("dmcf_b200 runs on CUDA devices only (no CPU fallback)")
        self.cfg = dict(kwargs, main_log_dir=main_log_dir, split=split)
        self.timing = []

    def step(self, inputs):
        """One model call on one sample ``[pos, vel, acc|None, feats|None, box, box_normals]`` -> the next sample
        (the body of run_inference, pipelines/simulator.py:68-70)."""
        pos, vel = self.model(inputs, training=False)
        slab = getattr(self.model, "slab", None)
        if slab is not None and slab.world > 1:
            # hand particles that crossed a slab face to the neighbouring rank (per-particle acc travels with them)
            if inputs[2] is not None:
                pos, vel, acc = slab.migrate(pos, vel, inputs[2])
                return [pos, vel, acc] + list(inputs[3:])
            pos, vel = slab.migrate(pos, vel)
        return [pos, vel] + list(inputs[2:])

    @torch.no_grad()
    def run_inference(self, inputs):
        """pipelines/simulator.py:57-71: list of samples in, list of advanced samples out."""
        return [self.step(sample) for sample in inputs]

    @torch.no_grad()
    def run_rollout(self, inputs, timesteps=2):
        """pipelines/simulator.py:73-109.  ``inputs`` is a list of dicts with 'pos','vel','grav','box','box_normals'
        arrays whose first axis is time (frame 0 is used)."""
        dev = self.device
        samples = [[_to_dev(d["pos"][0], dev), _to_dev(d["vel"][0], dev),
                    _to_dev(d["grav"][0], dev) if d.get("grav") is not None and d["grav"][0] is not None else None, None,
                    _to_dev(d["box"][0], dev), _to_dev(d["box_normals"][0], dev)] for d in inputs]
        results = [[s] for s in samples]
        self.timing = []
        for _ in range(timesteps - 1):
            torch.cuda.synchronize(dev)
            start = time.time()
            for i in range(len(samples)):
                samples[i] = self.run_inference(samples[i:i + 1])[0]
            torch.cuda.synchronize(dev)
            self.timing.append(time.time() - start)
            for i in range(len(samples)):
                results[i].append(samples[i])
        if self.timing:
            log.info("Average runtime: %.05f" % (np.mean(self.timing) / max(len(samples), 1)))
        return results

    def load_ckpt(self, ckpt_path=None, is_resume=True):
        """Restores model weights from a TF2 checkpoint prefix or directory; returns the epoch parsed from the name."""
        import re
        if ckpt_path is None:
            log.info("Initializing from scratch.")
            return 0
        ckpt, prefix = load_checkpoint(ckpt_path, return_prefix=True)
        weights = model_weights(ckpt)
        missing = self.model.load_weights(weights, device=self.device)
        if missing:
            log.warning("layers without checkpoint weights: %s", missing)
        log.info("Restored from {}".format(ckpt_path))
        nums = re.findall(r"\d+", str(prefix).split("/")[-1])  # epoch from the bundle that was restored ("ckpt-12")
        return int(nums[-1]) if nums else 0
