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


class _PlannedStep:
    """State of the sync-free step of one (particle count, boundary) signature: the plan of its data-dependent sizes, the
    pinned copy of its overflow flags, and (graph mode) the captured CUDA graph with its static input / output buffers."""

    def __init__(self, sig):
        self.sig = sig
        self.plan = None
        self.replays = 0
        self.graph = None
        self.graph_failed = False
        self.static_in = None
        self.static_out = None
        self.graph_launches = 0  # kernels of libdmcf_b200 inside one replay of the graph
        self.graph_shape = None
        self.flags_host = None
        self.event = None


class Simulator:
    """Pipeline for the trainable simulator (inference only).

    ``step_mode`` selects how ``step`` drives the model (the reference compiles its step with tf.function,
    pipelines/simulator.py:57):
      * ``"eager"``   every data-dependent size (cell-grid extent, pairs per neighbour list, culled boundary rows, lattice
                      points per scale) is read back from the device when it is needed: ~10-60 host syncs per step;
      * ``"planned"`` the first step of a scene measures those sizes (eager), later steps replay the plan on capacity-sized
                      buffers with device-side counts: NO host sync inside the step, one overflow-flag read after it;
      * ``"graph"``   (default) the planned step captured in a CUDA graph and replayed: one launch per step.
    A step whose plan overflows (or whose particles leave the planned cell grids) is detected by the flag read, recomputed
    exactly by a fresh measuring step and re-planned; results returned by ``step`` are always those of a valid step.
    Models that cannot be planned (layer-by-layer path, slab decomposition) run eagerly."""

    def __init__(self, model, dataset=None, name="Simulator", main_log_dir="./logs/", device="cuda", split="test",
                 step_mode="graph", **kwargs):
        self.model = model
        self.dataset = dataset
        self.name = name
        self.device = torch.device(device if device != "gpu" else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("dmcf_b200 runs on CUDA devices only (no CPU fallback)")
        if step_mode not in ("eager", "planned", "graph"):
            raise ValueError("step_mode must be 'eager', 'planned' or 'graph'")
        self.step_mode = step_mode
        self.cfg = dict(kwargs, main_log_dir=main_log_dir, split=split)
        self.timing = []
        self._planned = None
        self.replan_log = []  # (hard | soft, [(plan entry, kind), ...]) of every re-plan
        self.stats = {"measured": 0, "replayed": 0, "graph_replays": 0, "replans": 0, "captures": 0}

    # -- one step ---------------------------------------------------------------------------------------------------------
    def _plannable(self, inputs):
        m = self.model
        slab = getattr(m, "slab", None)
        if self.step_mode == "eager" or not getattr(m, "fused", False):
            return False
        if (slab is None or slab.world == 1) and inputs[0].shape[0] == 0:
            return False  # (under slab decomposition every rank must take the same path: empty slabs are planned as well)
        # the planned step is inference only (no autograd through capacity-sized buffers)
        return not any(p.requires_grad for p in m.parameters()) and not any(
            t is not None and t.requires_grad for t in inputs[:3])

    def step(self, inputs):
        """One model call on one sample ``[pos, vel, acc|None, feats|None, box, box_normals]`` -> the next sample
        (the body of run_inference, pipelines/simulator.py:68-70)."""
        if self._plannable(inputs):
            with torch.no_grad():
                out = self._step_planned(inputs)
            return list(out) + list(inputs[len(out):])
        pos, vel = self.model(inputs, training=False)
        slab = getattr(self.model, "slab", None)
        if slab is not None and slab.world > 1:
            # hand particles that crossed a slab face to the neighbouring rank (per-particle acc travels with them)
            if inputs[2] is not None:
                pos, vel, acc = slab.migrate(pos, vel, inputs[2])
                return [pos, vel, acc] + list(inputs[3:])
            pos, vel = slab.migrate(pos, vel)
        return [pos, vel] + list(inputs[2:])

    def _slab(self):
        slab = getattr(self.model, "slab", None)
        return slab if (slab is not None and slab.world > 1) else None

    def _run_model(self, inputs, plan, mode):
        """The model step (plus, under slab decomposition, the migration of particles that crossed a face) under ``plan``.
        Returns (pos, vel) or (pos, vel, acc)."""
        from . import ops
        plan.begin(mode)
        ops.set_plan(plan)
        try:
            pos, vel = self.model(inputs, training=False)
            slab = self._slab()
            if slab is None:
                return pos, vel
            if inputs[2] is not None:
                return slab.migrate(pos, vel, inputs[2])
            return slab.migrate(pos, vel)
        finally:
            ops.set_plan(None)

    def _step_planned(self, inputs):
        from . import ops
        pos, vel, acc, feats, box, bn = inputs
        slab = self._slab()
        # a slab's particle count changes with every migration: its plan is bounded by capacities, not by the input shape
        sig = (pos.shape[0] if slab is None else -1, acc is None, box.data_ptr(), box.shape[0], box._version, bn.data_ptr(),
               bn._version)
        st = self._planned
        if st is None or st.sig != sig:
            st = self._planned = _PlannedStep(sig)  # new scene / particle count (inflow): plan again
        if st.plan is None:
            # measuring step: the eager path (exact sizes, host syncs) with the sizes recorded; capacity-sized inputs (the state
            # a replayed slab step returned) are cut to their exact size first
            plan = ops.StepPlan(pos.device)
            inputs = [ops.trim(t) if (i < 3 and t is not None) else t for i, t in enumerate(inputs)]
            out = self._run_model(inputs, plan, "measure")
            st.plan, st.replays, st.graph = plan, 0, None
            self.stats["measured"] += 1
            return out
        if st.event is None:
            st.event = torch.cuda.Event()
        from . import ops as _o
        shape_key = (tuple(pos.shape), _o.count_of(pos) is not None)
        if st.graph is not None and st.graph_shape != shape_key:
            st.graph = None  # a slab's row capacity settles over the first steps of a rollout: capture again for the new shape
        graphable = slab is None or getattr(slab.transport, "graph_capturable", False)
        if self.step_mode == "graph" and graphable and st.graph is None and not st.graph_failed and st.replays >= 1:
            self._capture(st, inputs)
            st.graph_shape = shape_key
        if st.graph is not None:
            for dst, src in zip(st.static_in, (pos, vel, acc)):
                if dst is not None:
                    dst.copy_(src)
                    if _o.count_of(dst) is not None:
                        _o.count_of(dst).copy_(_o.count_of(src))
            st.graph.replay()
            out = tuple(self._clone_with_count(t) for t in st.static_out)
            self.stats["graph_replays"] += 1
            self.stats["graph_kernel_launches"] = self.stats.get("graph_kernel_launches", 0) + st.graph_launches
        else:
            out = self._run_model(inputs, st.plan, "replay")
            if st.flags_host is None or st.flags_host.shape != st.plan.flags.shape:
                st.flags_host = torch.zeros(st.plan.flags.shape, dtype=torch.int32).pin_memory()
            if slab is not None:  # every rank must reach the same verdict: message capacities are derived from the shared plan
                slab.all_reduce_max(st.plan.flags)
            st.flags_host.copy_(st.plan.flags, non_blocking=True)
        st.replays += 1
        self.stats["replayed"] += 1
        # the ONE read-back of the step: its overflow flags, after everything was enqueued
        st.event.record()
        st.event.synchronize()
        n = len(st.plan.entries)
        hard, soft = bool(st.flags_host[:n].any()), bool(st.flags_host[n:].any())
        if hard or soft:
            idx = torch.nonzero(st.flags_host).flatten().tolist()
            self.replan_log.append(("hard" if hard else "soft", [(i % n, st.plan.entries[i % n]["kind"]) for i in idx]))
        if hard:  # something outgrew its capacity: the results are invalid -> exact step + new plan
            log.info("step plan overflowed (%s): re-planning", self.replan_log[-1][1])
            st.plan, st.graph = None, None
            self.stats["replans"] += 1
            self.stats["replans_hard"] = self.stats.get("replans_hard", 0) + 1
            return self._step_planned(inputs)
        if soft:  # particles left a planned cell grid (results are fine, the grid is just no longer tight): plan again next step
            st.plan, st.graph = None, None
            self.stats["replans"] += 1
        return out

    @staticmethod
    def _clone_with_count(t):
        from . import ops
        c = t.clone()
        if ops.count_of(t) is not None:
            ops.with_count(c, ops.count_of(t).clone())
        return c

    def _capture(self, st, inputs):
        """Captures the replaying step on static input buffers.  Any failure (an op that cannot be captured) leaves the
        simulator in planned mode."""
        from . import ops
        pos, vel, acc, feats, box, bn = inputs
        saved_profile, ops.PROFILE = ops.PROFILE, None  # per-launch CUDA events cannot be recorded into a capture
        try:
            st.static_in = [torch.empty_like(pos), torch.empty_like(vel), None if acc is None else torch.empty_like(acc)]
            cnt = ops.count_of(pos)
            static_cnt = None if cnt is None else cnt.clone()
            for dst, src in zip(st.static_in, (pos, vel, acc)):
                if dst is not None:
                    dst.copy_(src)
                    if static_cnt is not None:
                        ops.with_count(dst, static_cnt)
            if st.flags_host is None or st.flags_host.shape != st.plan.flags.shape:
                st.flags_host = torch.zeros(st.plan.flags.shape, dtype=torch.int32).pin_memory()
            torch.cuda.synchronize(pos.device)
            g = torch.cuda.CUDAGraph()
            launches0 = ops.launch_count()
            slab = self._slab()
            # thread_local: other threads (NCCL's watchdog, the rank threads of an in-process transport) keep using CUDA freely
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                out = self._run_model([st.static_in[0], st.static_in[1], st.static_in[2], feats, box, bn], st.plan, "replay")
                if slab is not None:
                    slab.all_reduce_max(st.plan.flags)
                st.flags_host.copy_(st.plan.flags, non_blocking=True)
            st.graph, st.static_out, st.graph_launches = g, out, ops.launch_count() - launches0
            self.stats["captures"] += 1
        except Exception as e:  # noqa: BLE001
            log.warning("CUDA graph capture of the step failed (%s: %s); staying in planned mode", type(e).__name__, e)
            st.graph, st.graph_failed = None, True
            torch.cuda.synchronize(pos.device)
        finally:
            ops.PROFILE = saved_profile

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
