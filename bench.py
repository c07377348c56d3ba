"""bench.py -- particles*steps/s of DMCF's per-step hot path on B200 (BASELINE.json north_star).

Workload (config 4 of BASELINE.json): synthetic 3-D box of jittered-lattice fluid particles (spacing 0.05) inside wall
particles, single-scale SymNet = input convs -> 3x ContinuousConv(4x4x4, ->32) + Dense -> antisymmetric
ContinuousConv(6x6x6, 32->3), seeded random weights.  One "step" = one full model step (integrate, cull, neighbour
search, conv stack, correction) on the resident scene; every step starts from the same state so the work per step is
fixed.  N > 1 (torchrun, one rank per GPU): STRONG scaling of that one ~1 M-particle scene -- slabs of n_side / N lattice layers
along x with halo exchange -- is the headline (`"scaling": "strong"`); the weak-scaling run (every rank owns an n_side^3 slab
of an N times longer box) rides along as the extra key `weak`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n-side 100] [--impl ours|reference] [--scaling strong|weak|both]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference path (oracle O32 timed build, all
host threads, thread count set explicitly) on the same full scene; steps are cut to a 150 s budget.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particles_steps_per_sec"
UNIT = "particles*steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.lines, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                mx.append(cmax)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # timed region shorter than one sample: fall back to everything we saw
            for ts, line in self.lines:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except (ValueError, IndexError):
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(n_side, seed):
    from dmcf_b200 import scenes
    scene = scenes.lattice_scene((n_side, n_side, n_side), dx=0.05, jitter=0.2, vel_sigma=0.1, seed=seed)
    return scene, scenes.c4_model_cfg()


def oracle_weights(model):
    w = model.state_arrays()
    for name, layer in model.named_layers().items():
        for a in getattr(layer, "_aliases", []):
            if name + "/kernel" in w:
                w[a + "/kernel"] = w[name + "/kernel"]
                if name + "/bias" in w:
                    w[a + "/bias"] = w[name + "/bias"]
    return w


def host_threads():
    """Host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, so the CPU arm sets its OpenMP
    thread count explicitly (and prints it) instead of inheriting that."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(n_side_sample, steps, warmup, seed=0, budget_s=None):
    """Times the O32 restatement of the reference step (one search per conv, two-pass ASCC, separate Dense / relu / add) on the
    host cores: the TIMED build of oracle/o32.c (FMA contraction allowed, register-blocked patch x filter product), all host
    threads.  ``budget_s``: stop adding timed steps once the budget is used (at least one step is always timed)."""
    from dmcf_b200 import config
    from oracle import o32
    o32.use_timed_build(True)
    o32.set_num_threads(host_threads())
    scene, cfg = build_workload(n_side_sample, seed)
    model = config.build_model(cfg)
    model.init_weights(seed=0, device="cpu", scale=0.1)
    ref = o32.ModelO32(cfg, oracle_weights(model))
    n = scene["pos"].shape[0]
    times = []
    t_start = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        ref(scene["pos"], scene["vel"], None, scene["box"], scene["box_normals"])
        if i >= warmup:
            times.append(time.time() - t0)
        if budget_s is not None and times and time.time() - t_start + times[-1] > budget_s:
            break
    sec = float(np.mean(times))
    return {"value": n / sec, "unit": UNIT, "cores": o32.num_threads(), "kind": "port",
            "sample": f"{n_side_sample}^3 = {n} fluid particles + {scene['box'].shape[0]} wall particles, same net/seed, "
                      f"{len(times)} step(s) after {warmup} warm-up, {sec:.2f} s/step (oracle O32, timed build: C/OpenMP float32 "
                      f"restatement with FMA, one neighbour search per conv like the reference; {o32.num_threads()} OpenMP threads "
                      "set explicitly)"}, sec, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warm = 1 if args.warmup > 0 else 0
    base, sec, n_timed = cpu_baseline(args.cpu_n_side or args.n_side, max(1, args.steps), warm, budget_s=150.0)
    full = (args.cpu_n_side or args.n_side) == args.n_side
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": n_timed,
            "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(args.n_side) if full else
                       f"C4 synthetic 3-D box, single-scale SymNet (ASCC+CConv stack); CPU sample {args.cpu_n_side}^3 particles "
                       f"(bounded sub-volume of the {args.n_side}^3 workload)",
                       "same_config_as_gpu_arm": full,
                       "steps_note": f"{n_timed} of the requested {args.steps} steps fit the 150 s budget of the CPU arm"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_name(n_side):
    return (f"C4: ONE synthetic 3-D box of {n_side}^3 = {n_side ** 3} fluid particles (spacing 0.05, jittered lattice) inside wall "
            "particles, single-scale SymNet (input convs, 3x CConv 4x4x4 ->32 + Dense, ASCC 6x6x6 32->3), r=0.1, seeded random "
            "weights")


def ncu_traffic(top):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/roofline_traffic.json, written by hand from profiles/*_full.md); None if not captured."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    key = f"{top['kernel']}:{'x'.join(str(v) for v in top['filter'])}:{top['cin']}->{top['cout']}"
    return json.load(open(p)).get(key)


def conv_algorithmic_bytes(rec):
    """N_in(12+4Cin) + N_out(12+4Cout) + 4 K Cin Cout (+ residual read) (+ CSR since we consume one)."""
    b = rec["n_inp"] * (12 + 4 * rec["cin"]) + rec["n_out"] * (12 + 4 * rec["cout"]) + 4 * rec["rows"] * rec["cout"]
    if rec["residual"]:
        b += rec["n_out"] * 4 * rec["cout"]
    b += 4 * rec["pairs"] + 8 * (rec["n_out"] + 1)
    return b


def hbm_op_breakdown(recs, step_ms_total, peak):
    """The HBM-bound ops around the convs (cell list, neighbour search, pair geometry): live CUDA-event time per launch in
    the timed region against their algorithmic bytes (B_frs = 12 (N_in + N_out) + 4 P + 8 (N_out + 1); cell list
    12 N read + 20 N written + 4 per cell; pair geometry 4 P + 12 (N_in + N_out) + 8 (N_out + 1) read, 36 P written)."""
    def nbytes(r):
        if r["kind"] == "cell_list":
            return 32 * r["n_points"] + 4 * r["n_cells"]
        if r["kind"] == "frs_count":
            return 12 * (r["n_points"] + r["n_queries"]) + 4 * r["n_queries"] + 8 * (r["n_queries"] + 1)
        if r["kind"] == "frs_fill":
            return 12 * (r["n_points"] + r["n_queries"]) + (8 if r["distances"] else 4) * r["pairs"] + 8 * (r["n_queries"] + 1)
        if r["kind"] == "pair_records":
            return 40 * r["pairs"] + 12 * (r["n_inp"] + r["n_out"]) + 8 * (r["n_out"] + 1)
        return 0
    names = {"cell_list": "dmcf_grid_build (k_cell_hist/scan/scatter/sort/gather_pos)", "frs_count": "k_frs<count> + row_splits scan",
             "frs_fill": "k_frs<fill>", "pair_records": "k_cconv_prepare"}
    groups = {}
    for r in recs:
        g = groups.setdefault(r["kind"], {"ms": 0.0, "n": 0, "bytes": 0})
        g["ms"] += r["start"].elapsed_time(r["end"])
        g["n"] += 1
        g["bytes"] += nbytes(r)
    out = []
    for kind, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
        gbps = g["bytes"] / 1e9 / (g["ms"] * 1e-3) if g["ms"] > 0 else 0.0
        out.append({"kernel": names.get(kind, kind), "launches": g["n"], "avg_ms": round(g["ms"] / g["n"], 4),
                    "share_of_step": round(g["ms"] / step_ms_total, 4), "algorithmic_GB_per_launch": round(g["bytes"] / g["n"] / 1e9, 4),
                    "GBps": round(gbps, 1), "frac_of_hbm_peak": round(gbps / peak, 4)})
    return out


def conv_flops(rec):
    """P(F_map + 16 Cin) + 2 N_out K Cin Cout."""
    return rec["pairs"] * (60 + 16 * rec["cin"]) + 2 * rec["n_out"] * rec["rows"] * rec["cout"]


def timed_steps(sim, sample, steps, warmup, barrier, local_rank, profile=True):
    """W untimed + exactly K timed steps on the resident sample, bracketed by barrier + synchronize; CUDA events on the launching
    stream; clocks sampled during the timed region.  Returns (ms of the K steps on this rank, conv / hbm-op records, launches,
    clocks, last output)."""
    import torch
    from dmcf_b200 import ops
    with torch.no_grad():
        for _ in range(warmup):
            out = sim.step(sample)
        barrier()
        ops.PROFILE = [] if profile else None
        launches0 = ops.launch_count() + sim.stats.get("graph_kernel_launches", 0)
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.25)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        ev0.record()
        for _ in range(steps):
            out = sim.step(sample)
        ev1.record()
        barrier()
        w1 = time.time()
        clocks = sampler.stop(w0, w1)
        ms = ev0.elapsed_time(ev1)
        # kernels launched one by one plus the kernels inside every replayed CUDA graph (counted when the graph was captured)
        launches = ops.launch_count() + sim.stats.get("graph_kernel_launches", 0) - launches0
        prof, ops.PROFILE = (ops.PROFILE or []), None
    return ms, prof, launches, clocks, out


def e2e_steps_timed(sim, scene, box_d, bn_d, steps, barrier, dev, return_host_results=False):
    """The same metric through the public API with HOST buffers: every step copies this rank's pos / vel from pinned host memory,
    runs Simulator.step and copies the advanced pos / vel back to pinned host memory, all inside the timed region.  The copies run
    on two side streams like a data loader would run them: the inputs of step k + 1 travel while step k computes, the results of
    step k while step k + 1 computes (every step still waits for ITS inputs and every result is copied out before the clock stops)."""
    import torch
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).pin_memory()
    h_pos, h_vel = pin(scene["pos"]), pin(scene["vel"])
    o_pos, o_vel = torch.empty_like(h_pos).pin_memory(), torch.empty_like(h_vel).pin_memory()
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    # two device input buffers: the copy for step k + 1 runs while step k computes, into the buffer step k - 1 has finished with
    # (no allocations on the side streams: a block freed there would wait for cross-stream events in the caching allocator)
    d_pos = [torch.empty(h_pos.shape, dtype=torch.float32, device=dev) for _ in range(2)]
    d_vel = [torch.empty(h_vel.shape, dtype=torch.float32, device=dev) for _ in range(2)]
    read_done = [None, None]

    def fetch(k):
        b = k & 1
        with torch.cuda.stream(s_in):
            if read_done[b] is not None:
                s_in.wait_event(read_done[b])
            d_pos[b].copy_(h_pos, non_blocking=True)
            d_vel[b].copy_(h_vel, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(s_in)
        return b, ev

    with torch.no_grad():
        def run(k_steps):
            nxt = fetch(0)
            for k in range(k_steps):
                b, ev = nxt
                main.wait_event(ev)
                if k + 1 < k_steps:
                    nxt = fetch(k + 1)
                res = sim.step([d_pos[b], d_vel[b], None, None, box_d, bn_d])
                done = torch.cuda.Event()
                done.record(main)
                read_done[b] = done
                n = min(res[0].shape[0], o_pos.shape[0])  # slabs: migration may change the row count by a few
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    o_pos[:n].copy_(res[0][:n], non_blocking=True)
                    o_vel[:n].copy_(res[1][:n], non_blocking=True)
                res[0].record_stream(s_out)
                res[1].record_stream(s_out)
            main.wait_stream(s_out)  # the last results are on the host before the region ends
        run(1)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(steps)
        e1.record()
        barrier()
    if return_host_results:
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), int(h_pos.numel() * 4 + h_vel.numel() * 4), o_pos, o_vel
    return e0.elapsed_time(e1), int(h_pos.numel() * 4 + h_vel.numel() * 4)


def run_c5(args):
    """BASELINE config 5: synthetic 3-D open box of N x n_side^3 particles (4 M at N = 8, n_side = 80), the full multi-scale
    Liquid3d net with the shipped checkpoint, slab partitioned, an EVOLVING rollout: every step advances the state of the one
    before, particles migrate between slabs, neighbour counts / lattices / culled walls change, plans are re-made when a
    capacity overflows.  K timed steps after W warm-up steps."""
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from dmcf_b200 import config, ops, scenes
    from dmcf_b200.simulator import Simulator
    from dmcf_b200.slab import SlabContext
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_side = args.n_side if args.n_side != 100 else 80
    scene, faces = scenes.slab_scene(n_side, rank, world, dx=0.05, jitter=0.2, vel_sigma=0.05, seed=2, open_top=True)
    z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_Liquid3d.npz"))
    weights = {k.replace("|", "/"): z[k] for k in z.files}
    # gravity off: the shipped checkpoint does not hold a synthetic lattice block against a synthetic floor (particles leak through
    # it and the scene explodes, scripts/dbg_stability.py); without gravity its corrections let the block relax and expand slowly
    # (|v| ~ 0.3-0.7 m/s): a tame state that still changes every step -- neighbour counts, lattices, culled walls, slab ownership
    model = config.build_model(dict(scenes.liquid3d_model_cfg(), grav=0.0))
    assert model.load_weights(weights, device=dev) == []
    if world > 1:
        model.set_slab(SlabContext(faces, axis=0))
    sim = Simulator(model, device=f"cuda:{local_rank}", step_mode=args.step_mode)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    state = [t(scene["pos"]), t(scene["vel"]), None, None, t(scene["box"]), t(scene["box_normals"])]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def all_red(x, op):
        v = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=op)
        return float(v.item())

    warmup, steps = max(args.warmup, 3), args.steps
    torch.cuda.reset_peak_memory_stats(dev)
    with torch.no_grad():
        for _ in range(warmup):
            state = sim.step(state)
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.25)
        barrier()
        stats0 = dict(sim.stats)
        launches0 = ops.launch_count() + sim.stats.get("graph_kernel_launches", 0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        ev0.record()
        for _ in range(steps):
            state = sim.step(state)
        ev1.record()
        barrier()
        w1 = time.time()
        clocks = sampler.stop(w0, w1)
        ms = all_red(ev0.elapsed_time(ev1), dist.ReduceOp.MAX if world > 1 else None)
        launches = ops.launch_count() + sim.stats.get("graph_kernel_launches", 0) - launches0
        # end to end: the evolving state lives in pinned host memory between steps (capacity-sized rows, count read back)
        e2e_steps = max(3, min(steps, 10))
        pos_h = ops.trim(state[0]).cpu().pin_memory()
        vel_h = ops.trim(state[1]).cpu().pin_memory()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        copied = 0
        for _ in range(e2e_steps):
            st_in = [pos_h.to(dev, non_blocking=True), vel_h.to(dev, non_blocking=True)] + state[2:]
            out = sim.step(st_in)
            p_out, v_out = ops.trim(out[0]), ops.trim(out[1])
            pos_h = torch.empty(p_out.shape, dtype=torch.float32).pin_memory()
            vel_h = torch.empty(v_out.shape, dtype=torch.float32).pin_memory()
            pos_h.copy_(p_out, non_blocking=True)
            vel_h.copy_(v_out, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            copied += 2 * pos_h.numel() * 4
        e1.record()
        barrier()
        e_ms = all_red(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None)
    n_now = int(all_red(float(ops.trim(state[0]).shape[0]), dist.ReduceOp.SUM if world > 1 else None))
    n_total = int(all_red(float(scene["pos"].shape[0]), dist.ReduceOp.SUM if world > 1 else None))
    finite = all_red(0.0 if bool(torch.isfinite(ops.trim(state[0])).all().item()) else 1.0, dist.ReduceOp.SUM if world > 1 else None) == 0.0
    peak_gb = all_red(torch.cuda.max_memory_allocated(dev) / 2 ** 30, dist.ReduceOp.MAX if world > 1 else None)
    copied_total = int(all_red(float(copied), dist.ReduceOp.SUM if world > 1 else None)) // e2e_steps
    if rank == 0:
        d = {k: sim.stats.get(k, 0) - stats0.get(k, 0) for k in sim.stats}
        line = {"metric": METRIC, "value": n_total * steps / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
                "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"C5: synthetic 3-D open box, {world} x {n_side}^3 = {n_total} fluid particles, full multi-scale "
                                       "Liquid3d net (3 scales, shipped checkpoint, gravity off), EVOLVING rollout (state advances every step, "
                                       "migration between slabs, re-planning on overflow)",
                           "parallelism": f"{world} spatial slabs along x, halos per scale and layer over NCCL" if world > 1 else "single GPU",
                           "step_mode": sim.step_mode, "step_stats_timed_region": d, "particles_after": n_now,
                           "replan_reasons_first5": [list(r) for r in sim.replan_log[:5]],
                           "l2": "working set >> 126 MB L2", "peak_memory_GB_per_gpu": round(peak_gb, 2)},
                "e2e": {"value": n_total * e2e_steps / (e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": copied_total // 2,
                        "d2h_bytes_per_step": copied_total // 2, "steps": e2e_steps,
                        "api": "Simulator.step; the evolving pos / vel travel host -> device and back through pinned memory every step"},
                "gpu_launches": int(launches), "clocks": clocks, "finite": bool(finite), "roofline": None, "cpu_baseline": None}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n-side", type=int, default=100, help="fluid lattice edge of the scene (100 -> 1M particles)")
    ap.add_argument("--cpu-n-side", type=int, default=0, help="edge of the CPU arm's scene (0 = the full --n-side scene)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="both", choices=["strong", "weak", "both"],
                    help="N > 1: strong = ONE n_side^3 scene split into N slabs (the headline, BASELINE config 4); weak = every rank "
                         "owns an n_side^3 slab of an N-times longer box (extra key `weak`)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--step-mode", default="graph", choices=["eager", "planned", "graph"],
                    help="how Simulator.step drives the model (dmcf_b200/simulator.py); slab runs (N > 1) use 'planned' for 'graph'")
    ap.add_argument("--workload", default="c4", choices=["c4", "c5"],
                    help="c4 (default): the ~1 M-particle single-scale scene BASELINE's metric is quoted on; c5: N x 80^3 particles, full "
                         "multi-scale Liquid3d net, evolving rollout (BASELINE config 5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "c5":
        return run_c5(args)
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner at the first
    # communicator) are sent to stderr until the result line is written
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from dmcf_b200 import config, ops, scenes
    from dmcf_b200.simulator import Simulator
    from dmcf_b200.slab import SlabContext

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def all_max(x):
        v = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    def all_sum(x):
        v = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        return float(v.item())

    def make_sim(scene, faces):
        model = config.build_model(scenes.c4_model_cfg())
        model.init_weights(seed=0, device=dev, scale=0.1)
        if world > 1:
            model.set_slab(SlabContext(faces, axis=0))
        mode = args.step_mode
        sim = Simulator(model, device=f"cuda:{local_rank}", step_mode=mode)
        sample = [t(scene["pos"]), t(scene["vel"]), None, None, t(scene["box"]), t(scene["box_normals"])]
        return sim, sample

    # ---- the headline arm: ONE n_side^3 scene (strong scaling for N > 1: slabs of n_side / N lattice layers) ---------------
    full_scene, _ = build_workload(args.n_side, seed=0)
    n_total = full_scene["pos"].shape[0]
    n_wall_total = full_scene["box"].shape[0]
    if world > 1:
        scene, faces = scenes.slab_partition(full_scene, rank, world, args.n_side, dx=0.05)
    else:
        scene, faces = full_scene, None
    del full_scene
    main_scaling = "weak" if (world > 1 and args.scaling == "weak") else "strong"
    p_steps = 0
    weak = None
    if world > 1 and args.scaling in ("weak", "both"):
        w_scene, w_faces = scenes.slab_scene(args.n_side, rank, world, dx=0.05, jitter=0.2, vel_sigma=0.1, seed=0)
        sim_w, sample_w = make_sim(w_scene, w_faces)
        w_steps = args.steps if main_scaling == "weak" else max(3, min(args.steps, 10))
        ms_w, prof_w, launches_w, clocks_w, out_w = timed_steps(sim_w, sample_w, w_steps, warmup, barrier, local_rank,
                                                                profile=main_scaling == "weak")
        ms_w = all_max(ms_w)
        n_w = int(all_sum(w_scene["pos"].shape[0]))
        weak = {"value": n_w * w_steps / (ms_w * 1e-3), "unit": UNIT, "ms_per_step": ms_w / w_steps, "steps": w_steps,
                "particles_total": n_w, "particles_per_gpu": w_scene["pos"].shape[0],
                "workload": f"one {world}x{args.n_side} x {args.n_side} x {args.n_side} box, every rank owns an {args.n_side}^3 slab"}
        if main_scaling != "weak":
            del sim_w, sample_w, w_scene, out_w
            torch.cuda.empty_cache()

    if main_scaling == "weak":
        sim, sample, scene = sim_w, sample_w, w_scene
        ms, prof, launches, clocks, out = ms_w, prof_w, launches_w, clocks_w, out_w
        prof_region_ms = ms_w
        steps = w_steps
        n_total = weak["particles_total"]
        ms_max = ms
    else:
        sim, sample = make_sim(scene, faces)
        steps = args.steps
        replayed = args.step_mode != "eager"
        ms, prof, launches, clocks, out = timed_steps(sim, sample, steps, warmup, barrier, local_rank, profile=not replayed)
        ms_max = all_max(ms)
        if replayed:
            # per-kernel times for the roofline: the same step, plan and kernels launched one by one ("planned" mode) with CUDA
            # events around every launch, right after the headline region (events cannot be recorded inside a graph replay)
            sim_p = Simulator(sim.model, device=f"cuda:{local_rank}", step_mode="planned")
            p_steps = max(3, min(steps, 5))
            prof_region_ms, prof, _, _, _ = timed_steps(sim_p, sample, p_steps, 2, barrier, local_rank, profile=True)
        else:
            prof_region_ms = ms
    n_own = scene["pos"].shape[0]
    value = n_total * steps / (ms_max * 1e-3)

    # ---- roofline of the dominant kernel (live CUDA events around each conv launch in the timed region) -------
    groups = {}
    hbm_ops = [rec for rec in prof if "kind" in rec]
    prof = [rec for rec in prof if "kind" not in rec]
    for rec in prof:
        k = (rec["kernel_size"], rec["cin"], rec["cout"], rec["ascc"])
        g = groups.setdefault(k, {"ms": 0.0, "n": 0, "rec": rec})
        g["ms"] += rec["start"].elapsed_time(rec["end"])
        g["n"] += 1
    roofline, breakdown = None, []
    peak, peak_src = peaks()
    prof_ms_total = prof_region_ms  # the region the per-launch events were recorded in
    if groups:
        for k, g in sorted(groups.items(), key=lambda kv: -kv[1]["ms"]):
            avg_ms = g["ms"] / g["n"]
            gb = conv_algorithmic_bytes(g["rec"]) / 1e9
            breakdown.append({"kernel": g["rec"]["kernel"], "filter": list(k[0]), "cin": k[1], "cout": k[2], "ascc": bool(k[3]),
                              "launches": g["n"], "avg_ms": round(avg_ms, 4), "share_of_step": round(g["ms"] / prof_ms_total, 4),
                              "algorithmic_GB": round(gb, 4), "GBps": round(gb / (avg_ms * 1e-3), 1),
                              "fp32_TFLOPs": round(conv_flops(g["rec"]) / (avg_ms * 1e-3) / 1e12, 2)})
        top = breakdown[0]
        traffic = ncu_traffic(top) if world == 1 else None  # the capture is of the 1-GPU launch (1.06 M out points)
        roofline = {"bound": "hbm", "achieved": top["GBps"], "peak": peak, "unit": "GB/s", "frac": round(top["GBps"] / peak, 4),
                    "traffic": traffic, "kernel": f"{top['kernel']} filter {top['filter']} {top['cin']}->{top['cout']}" + (" ascc" if top["ascc"] else ""),
                    "peak_source": peak_src, "avg_launch_ms": top["avg_ms"], "share_of_step": top["share_of_step"],
                    "fp32_tflops": top["fp32_TFLOPs"], "fp32_simt_peak_tflops": 74.0,
                    "timed_in": ("per-launch CUDA events over %d steps launched kernel by kernel (step_mode planned: same plan, buffers and "
                                 "kernels as the graph) right after the headline region" % p_steps) if p_steps
                    else "per-launch CUDA events inside the headline timed region",
                    "note": "wide CConv layers are fp32-FLOP bound (SURVEY 8d): HBM fraction reported as BASELINE asks, "
                            "fp32 TFLOP/s beside it" + ("; rank 0's launches (its slab)" if world > 1 else "")}
        rec0 = next(g["rec"] for g in groups.values() if g["rec"]["kernel"] == top["kernel"] and g["rec"]["cin"] == top["cin"]
                    and g["rec"]["cout"] == top["cout"])
        if top["kernel"] == "k_cconv_lean" and top["cout"] == 32 and top["cin"] % 8 == 0 and top["cin"] > 8:
            # phase 2 of this kernel runs on the tensor cores (mma.sync m16n8k8 tf32, three passes for float32 parity): its floor
            # at the measured peak of that path (scripts/mma_sync_probe.cu: 511 MAC per clock and SM) against the launch
            macs = 3.0 * rec0["n_out"] * rec0["rows"] * rec0["cout"]
            floor_ms = macs / (511.0 * 148 * 1.965e9) * 1e3
            roofline["tensor_pipe"] = {"path": "mma.sync m16n8k8 tf32 x3 (HMMA.1688.F32.TF32)", "mac_per_launch": macs,
                                       "peak_mac_per_clk_sm": 511, "floor_ms": round(floor_ms, 3),
                                       "frac_of_launch": round(floor_ms / top["avg_ms"], 4)}

    hbm_kernels = hbm_op_breakdown(hbm_ops, prof_region_ms, peak)

    # ---- end-to-end arm: host buffers in, host buffers out, copies inside the timed region ----------------
    e2e_steps = max(3, min(steps, 10))
    e_ms, copy_bytes = e2e_steps_timed(sim, scene, sample[4], sample[5], e2e_steps, barrier, dev)
    e_ms = all_max(e_ms)
    e2e_value = n_total * e2e_steps / (e_ms * 1e-3)
    copy_bytes_total = int(all_sum(copy_bytes))

    # ---- sanity of the timed result (finite, particles moved) ------------------------------------------------
    ok = bool(all_sum(0.0 if bool(torch.isfinite(out[0]).all().item()) else 1.0) == 0.0)
    exchanged = None
    if world > 1:
        exchanged = int(all_sum(float(sim.model.slab.bytes_exchanged))) // max(1, warmup + steps + e2e_steps + 1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = cpu_baseline(args.cpu_n_side or args.n_side, 2, 1, budget_s=40.0)

    if rank == 0:
        layers = [args.n_side // world + (1 if k < args.n_side % world else 0) for k in range(world)]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms_max / steps, "higher_is_better": True, "scaling": main_scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": (workload_name(args.n_side) + f" ({n_total} fluid + {n_wall_total} wall particles in total)")
                       if main_scaling == "strong" else
                       f"C4 weak scaling: {world}x{args.n_side} x {args.n_side} x {args.n_side} box, {args.n_side}^3 fluid particles per GPU",
                       "parallelism": (f"{world} spatial slabs along x" + (f" of {layers} lattice layers" if main_scaling == "strong" else "")
                                       + ", position halo per step + feature halo per conv layer over NCCL send/recv, migration after the position update")
                       if world > 1 else "single GPU",
                       "l2": ("per-step working set (features 136 MB/layer + 140 MB neighbour list per 1 M particles) exceeds the 126 MB L2 "
                              "down to 8 slabs (17 MB/layer + 18 MB list each, but five layers + records > L2); no explicit flush"),
                       "state": "every step restarts from the same resident scene (fixed work per step)",
                       "step_mode": sim.step_mode + " (dmcf_b200/simulator.py)",
                       "step_stats": dict(sim.stats)},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": copy_bytes_total, "d2h_bytes_per_step": copy_bytes_total,
                    "steps": e2e_steps, "api": "Simulator.step on pinned host pos/vel, results copied back to pinned host; copies on side streams (inputs of step k+1 and results of step k-1 travel while step k computes)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "kernels": breakdown,
            "hbm_kernels": hbm_kernels, "cpu_baseline": cpu, "finite": ok,
            "particles_rank0": n_own, "halo_bytes_per_step": exchanged,
        }
        if weak is not None and main_scaling != "weak":
            line["weak"] = weak
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        # CUDA graphs that captured NCCL kernels are still alive: leave without tearing the communicator down under them
        # (destroy_process_group was seen to hang in that state); every rank has finished its work at this barrier
        barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
