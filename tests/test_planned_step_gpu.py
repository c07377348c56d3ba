"""The sync-free step ("no host sync, overflow flag read at step end"): Simulator.step in 'planned' and 'graph' mode
against the eager step on the same inputs, for every shipped config shape; zero host syncs inside a replayed step
(torch.cuda.set_sync_debug_mode); overflow / drift of a plan is detected and the step recomputed."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(__file__))


def build(cfg, cuda, weights=None, seed=0):
    from dmcf_b200 import config
    model = config.build_model(cfg)
    if weights is not None:
        assert model.load_weights(weights, device=cuda) == []
    else:
        model.init_weights(seed=seed, device=cuda, scale=0.1)
    return model


def sample_of(scene, cuda, acc=None):
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)
    return [t(scene["pos"]), t(scene["vel"]), t(acc), None, t(scene["box"]), t(scene["box_normals"])]


CASES = ["c4", "liquid3d", "wbc_sph_2d", "cconv"]


def make_case(name):
    import test_models_gpu as T
    from dmcf_b200 import scenes
    if name == "c4":
        return scenes.c4_model_cfg(), scenes.lattice_scene((14, 12, 10), dx=0.05, seed=1), None, None
    if name == "liquid3d":
        return (T.liquid3d_cfg(), scenes.lattice_scene((11, 10, 9), dx=0.05, seed=2, open_top=True),
                T.load_npz_weights("ckpt_Liquid3d.npz"), None)
    if name == "wbc_sph_2d":
        sc = scenes.hydrostatic_scene_2d(n=24)
        return T.wbc_cfg(), sc, T.load_npz_weights("ckpt_WBC-SPH.npz"), sc["acc"]
    return T.cconv_cfg(), scenes.lattice_scene((10, 9, 8), dx=0.05, seed=5), None, None


@pytest.mark.parametrize("mode", ["planned", "graph"])
@pytest.mark.parametrize("case", CASES)
def test_planned_and_graph_steps_match_the_eager_step(cuda, case, mode):
    from dmcf_b200.simulator import Simulator
    cfg, scene, weights, acc = make_case(case)
    model = build(cfg, cuda, weights)
    sample = sample_of(scene, cuda, acc)
    eager = Simulator(model, device="cuda", step_mode="eager")
    sim = Simulator(model, device="cuda", step_mode=mode)
    with torch.no_grad():
        ref = eager.step(sample)
        outs = [sim.step(sample) for _ in range(4)]  # measure, replay (eager launch), capture + graph replay, graph replay
    assert sim.stats["measured"] == 1 and sim.stats["replans"] == 0
    assert sim.stats["graph_replays"] == (2 if mode == "graph" else 0), sim.stats
    scale = float(ref[0].abs().max())
    for o in outs:
        # the measuring step IS the eager step; replays use padded cell grids (other summation order inside a row): rounding only
        assert float((o[0] - ref[0]).abs().max()) <= 2e-6 * max(scale, 1.0)
        assert float((o[1] - ref[1]).abs().max()) <= 2e-6 * max(scale, 1.0) / cfg["timestep"] * 2
    assert torch.equal(outs[0][0], ref[0])
    assert torch.equal(outs[2][0], outs[3][0])  # graph replays are deterministic


def test_replayed_step_makes_no_host_sync(cuda):
    """torch's sync debug mode raises on every synchronizing call torch makes (item / cpu / nonzero / boolean masks ...): a
    replayed step (planned or graph) must pass under it.  The step's one read-back -- its overflow flags, after the step -- is an
    event wait on a pinned buffer and is not a stream synchronisation."""
    import test_models_gpu as T
    from dmcf_b200 import scenes
    from dmcf_b200.simulator import Simulator
    for cfg, scene, weights in ((scenes.c4_model_cfg(), scenes.lattice_scene((14, 12, 10), dx=0.05, seed=1), None),
                                (T.liquid3d_cfg(), scenes.lattice_scene((11, 10, 9), dx=0.05, seed=2, open_top=True),
                                 T.load_npz_weights("ckpt_Liquid3d.npz"))):
        model = build(cfg, cuda, weights)
        sample = sample_of(scene, cuda)
        for mode in ("planned", "graph"):
            sim = Simulator(model, device="cuda", step_mode=mode)
            with torch.no_grad():
                for _ in range(3):
                    sim.step(sample)
                torch.cuda.synchronize()
                torch.cuda.set_sync_debug_mode("error")
                try:
                    for _ in range(3):
                        out = sim.step(sample)
                finally:
                    torch.cuda.set_sync_debug_mode("default")
            assert sim.stats["replans"] == 0
            assert bool(torch.isfinite(out[0]).all())
            if mode == "graph":
                assert sim.stats["graph_replays"] >= 3


def test_overflowing_plan_is_detected_and_the_step_recomputed(cuda):
    """Shrink the planned capacities behind the simulator's back: the replay truncates, the flag read finds it, the step is
    recomputed exactly and re-planned."""
    from dmcf_b200 import scenes
    from dmcf_b200.simulator import Simulator
    model = build(scenes.c4_model_cfg(), cuda)
    sample = sample_of(scenes.lattice_scene((14, 12, 10), dx=0.05, seed=1), cuda)
    sim = Simulator(model, device="cuda", step_mode="planned")
    with torch.no_grad():
        ref = sim.step(sample)
        sim.step(sample)
        for e in sim._planned.plan.entries:
            if e["kind"] == "pairs":
                e["total"] = int(e["total"] * 0.5)
        out = sim.step(sample)
    assert sim.stats["replans"] == 1 and sim.stats["measured"] == 2
    assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])


def test_graph_rollout_tracks_the_eager_rollout_with_evolving_state(cuda):
    """A short rollout with evolving particle state (neighbour counts, culled boundary rows and cell grids change every step):
    graph mode == eager mode step by step within rounding growth, with re-plans allowed."""
    import test_models_gpu as T
    from dmcf_b200 import scenes
    from dmcf_b200.simulator import Simulator
    model = build(T.liquid3d_cfg(), cuda, T.load_npz_weights("ckpt_Liquid3d.npz"))
    scene = scenes.lattice_scene((10, 12, 10), dx=0.05, seed=3, open_top=True)
    a = sample_of(scene, cuda)
    b = [t.clone() if t is not None else None for t in a]
    eager = Simulator(model, device="cuda", step_mode="eager")
    graph = Simulator(model, device="cuda", step_mode="graph")
    with torch.no_grad():
        for i in range(12):
            a, b = eager.step(a), graph.step(b)
            assert float((a[0] - b[0]).abs().max()) <= 1e-5 * (i + 1), i
    assert graph.stats["graph_replays"] >= 5, graph.stats


def test_bench_end_to_end_loop_returns_the_step_result(cuda):
    """bench.py's `e2e` loop moves the inputs and results of neighbouring steps on side streams while a step computes: what
    arrives in the pinned host buffers must be the result of Simulator.step on the same inputs."""
    import bench
    from dmcf_b200 import scenes
    from dmcf_b200.simulator import Simulator
    cfg, scene = scenes.c4_model_cfg(), scenes.lattice_scene((20, 16, 12), dx=0.05, seed=4)
    model = build(cfg, cuda)
    sample = sample_of(scene, cuda)
    sim = Simulator(model, device="cuda", step_mode="graph")
    with torch.no_grad():
        for _ in range(3):
            ref = sim.step(sample)
    _, nbytes, o_pos, o_vel = bench.e2e_steps_timed(sim, scene, sample[4], sample[5], 4, lambda: torch.cuda.synchronize(), cuda,
                                                    return_host_results=True)
    assert nbytes == scene["pos"].size * 4 * 2
    assert torch.equal(o_pos, ref[0].cpu()) and torch.equal(o_vel, ref[1].cpu())
