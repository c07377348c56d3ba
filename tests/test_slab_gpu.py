"""2-GPU slab run == 1-GPU run (needs >= 2 CUDA devices: `gpurun --gpus 2 -- pytest tests/test_slab_gpu.py -m gpu`)."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def global_scene():
    from dmcf_b200 import scenes
    return scenes.lattice_scene((28, 10, 10), dx=0.05, seed=21)


def model_cfg(kind):
    from dmcf_b200 import scenes
    if kind == "c4":
        return scenes.c4_model_cfg(), None
    import test_models_gpu as T
    return T.liquid3d_cfg(), T.load_npz_weights("ckpt_Liquid3d.npz")


def run_rank(rank, dev, slab, kind):
    """One rank's share of the scene through its own model + Simulator; returns what the comparison needs."""
    from dmcf_b200 import config
    from dmcf_b200.simulator import Simulator
    sc = global_scene()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    pos, box = t(sc["pos"]), t(sc["box"])
    own_f, own_b = slab.owned_mask(pos), slab.owned_mask(box)
    ids = torch.nonzero(own_f).flatten()
    if os.path.join(ROOT, "tests") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
    cfg, weights = model_cfg(kind)
    model = config.build_model(cfg)
    if weights is None:
        model.init_weights(seed=0, device=dev, scale=0.1)
    else:
        model.load_weights(weights, device=dev)
    model.set_slab(slab)
    Simulator(model, device=str(dev))
    sample = [pos[own_f], t(sc["vel"])[own_f], None, None, box[own_b], t(sc["box_normals"])[own_b]]
    with torch.no_grad():
        p1, v1 = model(sample)
    net = model.net_out[: int(own_f.sum())].cpu().numpy()
    return dict(ids=ids.cpu().numpy(), pos=p1.cpu().numpy(), vel=v1.cpu().numpy(), net=net,
                net_sum=model.net_out.double().sum(0).cpu().numpy(), net_abs=model.net_out.double().abs().sum(0).cpu().numpy(),
                bytes=slab.bytes_exchanged)


def worker(rank, world, port, out_dir, kind):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from dmcf_b200.slab import SlabContext
    faces = SlabContext.uniform_faces(0.0, 28 * 0.05, world)
    res = run_rank(rank, dev, SlabContext(faces, axis=0), kind)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **res)
    dist.barrier()
    dist.destroy_process_group()


def undivided_reference(kind, dev):
    from dmcf_b200 import config
    sc = global_scene()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    cfg, weights = model_cfg(kind)
    model = config.build_model(cfg)
    if weights is None:
        model.init_weights(seed=0, device=dev, scale=0.1)
    else:
        model.load_weights(weights, device=dev)
    with torch.no_grad():
        p_ref, v_ref = model([t(sc["pos"]), t(sc["vel"]), None, None, t(sc["box"]), t(sc["box_normals"])])
    return p_ref.cpu().numpy(), v_ref.cpu().numpy()


def check_ranks(ranks, p_ref, v_ref, kind):
    seen = np.zeros(len(p_ref), bool)
    tot_sum, tot_abs = 0.0, 0.0
    for r, z in enumerate(ranks):
        ids = z["ids"]
        assert not seen[ids].any()
        seen[ids] = True
        assert np.abs(z["pos"] - p_ref[ids]).max() <= (2e-6 if kind == "c4" else 1e-5), r
        assert np.abs(z["vel"] - v_ref[ids]).max() <= (2e-6 if kind == "c4" else 1e-5) / 0.02 * 2, r
        tot_sum, tot_abs = tot_sum + z["net_sum"], tot_abs + z["net_abs"]
        assert int(z["bytes"]) > 0
    assert seen.all()
    # momentum conservation across the slabs (fluid + boundary rows of all ranks)
    assert np.all(np.abs(tot_sum) <= 2e-5 * tot_abs + 1e-6)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["c4", "liquid3d_multiscale"])
def test_slabs_on_one_gpu_match_undivided(kind, world):
    """The same slab code (ownership, halo plans, per-layer feature halos, global bbox / mean reductions) with the ranks as
    threads of ONE process on ONE GPU and an in-process transport instead of NCCL: runs on a 1-GPU box."""
    import threading
    from dmcf_b200.slab import LocalTransport, SlabContext
    dev = torch.device("cuda:0")
    tr = LocalTransport(world)
    faces = SlabContext.uniform_faces(0.0, 28 * 0.05, world)
    out, errs = [None] * world, []

    def body(rank):
        try:
            torch.cuda.set_device(0)
            out[rank] = run_rank(rank, dev, SlabContext(faces, axis=0, rank=rank, world_size=world, transport=tr), kind)
        except BaseException as e:  # noqa: BLE001 -- re-raised in the main thread
            errs.append(e)
            tr._barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=500)
    if errs:
        raise errs[0]
    p_ref, v_ref = undivided_reference(kind, dev)
    check_ranks(out, p_ref, v_ref, kind)


def run_rank_planned(rank, dev, slab, kind, steps):
    """The rank's share through Simulator.step in 'planned' mode (sync-free replays of a measured plan, migration included),
    ``steps`` times from the same state; returns the trimmed outputs of the last step and the simulator's statistics."""
    from dmcf_b200 import config, ops
    from dmcf_b200.simulator import Simulator
    sc = global_scene()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    pos, box = t(sc["pos"]), t(sc["box"])
    own_f, own_b = slab.owned_mask(pos), slab.owned_mask(box)
    if os.path.join(ROOT, "tests") not in sys.path:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
    cfg, weights = model_cfg(kind)
    model = config.build_model(cfg)
    if weights is None:
        model.init_weights(seed=0, device=dev, scale=0.1)
    else:
        model.load_weights(weights, device=dev)
    model.set_slab(slab)
    sim = Simulator(model, device=str(dev), step_mode="planned")
    sample = [pos[own_f], t(sc["vel"])[own_f], None, None, box[own_b], t(sc["box_normals"])[own_b]]
    with torch.no_grad():
        for _ in range(steps):
            out = sim.step(sample)
    return dict(pos=ops.trim(out[0]).cpu().numpy(), vel=ops.trim(out[1]).cpu().numpy(), stats=dict(sim.stats),
                bytes=slab.bytes_exchanged)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("kind", ["c4", "liquid3d_multiscale"])
def test_planned_slabs_on_one_gpu_match_undivided(kind, world):
    """Sync-free slab steps (capacity-sized halo / migration messages, device-side counts, collective overflow verdict): after
    a measuring step and two replays the union of the ranks' particles (post migration) is the undivided step's result."""
    import threading
    from dmcf_b200.slab import LocalTransport, SlabContext
    dev = torch.device("cuda:0")
    tr = LocalTransport(world)
    faces = SlabContext.uniform_faces(0.0, 28 * 0.05, world)
    out, errs = [None] * world, []

    def body(rank):
        try:
            torch.cuda.set_device(0)
            out[rank] = run_rank_planned(rank, dev, SlabContext(faces, axis=0, rank=rank, world_size=world, transport=tr), kind, 3)
        except BaseException as e:  # noqa: BLE001 -- re-raised in the main thread
            errs.append(e)
            tr._barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=500)
    if errs:
        raise errs[0]
    p_ref, v_ref = undivided_reference(kind, dev)
    got_p = np.concatenate([z["pos"] for z in out])
    got_v = np.concatenate([z["vel"] for z in out])
    assert got_p.shape == p_ref.shape
    order_g = np.lexsort((got_p[:, 2], got_p[:, 1], got_p[:, 0]))
    order_r = np.lexsort((p_ref[:, 2], p_ref[:, 1], p_ref[:, 0]))
    tol = 2e-6 if kind == "c4" else 1e-5
    assert np.abs(got_p[order_g] - p_ref[order_r]).max() <= tol
    assert np.abs(got_v[order_g] - v_ref[order_r]).max() <= tol / 0.02 * 2
    for r, z in enumerate(out):  # every rank holds exactly the particles whose new position lies in its slab
        x = z["pos"][:, 0]
        assert np.all((x >= faces[r]) & (x < faces[r + 1]))
        assert z["stats"]["measured"] == 1 and z["stats"]["replayed"] == 2 and z["stats"]["replans"] == 0, z["stats"]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["c4", "liquid3d_multiscale"])
def test_two_gpu_slab_matches_single_gpu(tmp_path, kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from dmcf_b200 import config, scenes
    port = 29700 + os.getpid() % 2000 + (7 if kind == "c4" else 0)
    mp.spawn(worker, args=(2, port, str(tmp_path), kind), nprocs=2, join=True)
    p_ref, v_ref = undivided_reference(kind, torch.device("cuda:0"))
    check_ranks([dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(2)], p_ref, v_ref, kind)
