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


def worker(rank, world, port, out_dir, kind):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from dmcf_b200 import config, scenes
    from dmcf_b200.simulator import Simulator
    from dmcf_b200.slab import SlabContext
    sc = global_scene()
    faces = SlabContext.uniform_faces(0.0, 28 * 0.05, world)
    slab = SlabContext(faces, axis=0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    pos, box = t(sc["pos"]), t(sc["box"])
    own_f, own_b = slab.owned_mask(pos), slab.owned_mask(box)
    ids = torch.nonzero(own_f).flatten()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    cfg, weights = model_cfg(kind)
    model = config.build_model(cfg)
    if weights is None:
        model.init_weights(seed=0, device=dev, scale=0.1)
    else:
        model.load_weights(weights, device=dev)
    model.set_slab(slab)
    sim = Simulator(model, device=f"cuda:{rank}")
    sample = [pos[own_f], t(sc["vel"])[own_f], None, t(ids.float().cpu().numpy()[:, None]), box[own_b], t(sc["box_normals"])[own_b]]
    p1, v1 = model(sample[:3] + [None] + sample[4:])
    net = model.net_out[: int(own_f.sum())].cpu().numpy()
    n_own = model.net_out.shape[0]
    full_net_sum = model.net_out.double().sum(0).cpu().numpy()
    full_net_abs = model.net_out.double().abs().sum(0).cpu().numpy()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=ids.cpu().numpy(), pos=p1.cpu().numpy(), vel=v1.cpu().numpy(), net=net,
             net_sum=full_net_sum, net_abs=full_net_abs, bytes=slab.bytes_exchanged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("kind", ["c4", "liquid3d_multiscale"])
def test_two_gpu_slab_matches_single_gpu(tmp_path, kind):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from dmcf_b200 import config, scenes
    port = 29700 + os.getpid() % 2000 + (7 if kind == "c4" else 0)
    mp.spawn(worker, args=(2, port, str(tmp_path), kind), nprocs=2, join=True)
    dev = torch.device("cuda:0")
    sc = global_scene()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    cfg, weights = model_cfg(kind)
    model = config.build_model(cfg)
    if weights is None:
        model.init_weights(seed=0, device=dev, scale=0.1)
    else:
        model.load_weights(weights, device=dev)
    p_ref, v_ref = model([t(sc["pos"]), t(sc["vel"]), None, None, t(sc["box"]), t(sc["box_normals"])])
    p_ref, v_ref = p_ref.cpu().numpy(), v_ref.cpu().numpy()
    seen = np.zeros(len(p_ref), bool)
    tot_sum, tot_abs = 0.0, 0.0
    for r in range(2):
        z = np.load(tmp_path / f"rank{r}.npz")
        ids = z["ids"]
        seen[ids] = True
        assert np.abs(z["pos"] - p_ref[ids]).max() <= (2e-6 if kind == "c4" else 1e-5), r
        assert np.abs(z["vel"] - v_ref[ids]).max() <= (2e-6 if kind == "c4" else 1e-5) / 0.02 * 2, r
        tot_sum, tot_abs = tot_sum + z["net_sum"], tot_abs + z["net_abs"]
        assert int(z["bytes"]) > 0
    assert seen.all()
    # momentum conservation across the two slabs (fluid + boundary rows of both ranks)
    assert np.all(np.abs(tot_sum) <= 2e-5 * tot_abs + 1e-6)
