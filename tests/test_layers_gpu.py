"""Layer-API parity: the ContinuousConv / PointSampling classes and the lattice sampler with the reference's call
signatures (utils/convolutions.py:150-175, 277-286, 904-937; utils/tools/losses.py:136-181, 249-284)."""
import numpy as np
import pytest
import torch

from oracle import o64

pytestmark = pytest.mark.gpu


def close(got, ref, scale=1.0):
    ref = np.asarray(ref, np.float64)
    got = got.detach().cpu().numpy().astype(np.float64) if isinstance(got, torch.Tensor) else np.asarray(got, np.float64)
    tol = scale * (2e-5 * max(np.abs(ref).max(initial=0.0), 1e-30) + 1e-6)
    err = np.abs(got - ref).max(initial=0.0)
    assert err <= tol, f"{err:.3e} > {tol:.3e}"


@pytest.fixture
def cloud(cuda):
    rng = np.random.default_rng(5)
    pts = (rng.random((700, 3)) * 0.5).astype(np.float32)
    feats = rng.standard_normal((700, 6)).astype(np.float32)
    return rng, pts, feats, torch.from_numpy(pts).to(cuda), torch.from_numpy(feats).to(cuda)


def test_continuous_conv_layer_defaults_and_attributes(cuda, cloud):
    from dmcf_b200.convolutions import ContinuousConv
    rng, pts, feats, tp, tf = cloud
    conv = ContinuousConv(filters=5, kernel_size=[3, 3, 3], activation="relu")  # reference defaults otherwise
    out = conv(tf, tp, tp[:300], extents=0.25)
    assert tuple(conv.kernel.shape) == (3, 3, 3, 6, 5) and tuple(conv.bias.shape) == (5,) and conv.in_channels == 6
    assert float(conv.kernel.abs().max()) <= 0.05 and float(conv.bias.abs().sum()) == 0  # Keras 'uniform' / 'zeros'
    nns = conv.nns
    assert nns.neighbors_index.dtype == torch.int32 and nns.neighbors_row_splits.dtype == torch.int64
    assert nns.neighbors_distance.numel() == 0  # no window -> return_distances False (utils/convolutions.py:210)
    ref = o64.cconv_layer(feats, pts, pts[:300], np.float32(0.25), conv.kernel.cpu().numpy(), conv.bias.cpu().numpy(),
                          normalize=True, coordinate_mapping="ball_to_cube_radial")
    close(out, np.maximum(ref, 0))
    assert conv.compute_output_shape(None) == (None, 5)


def test_window_callable_and_user_neighbors_and_importance(cuda, cloud):
    from dmcf_b200 import ops
    from dmcf_b200.convolutions import ContinuousConv
    from dmcf_b200.losses import get_window_func
    rng, pts, feats, tp, tf = cloud
    imp = rng.random(700).astype(np.float32)
    win = lambda q: torch.clamp((1 - q) ** 3, 0, 1)  # the docstring's example window: NOT a WindowFunction -> unfused path
    a = ContinuousConv(4, [4, 4, 4], window_function=win, normalize=False, use_bias=False,
                       coordinate_mapping="ball_to_cube_volume_preserving", radius_search_ignore_query_points=True)
    b = ContinuousConv(4, [4, 4, 4], window_function=get_window_func("poly6"), normalize=False, use_bias=False,
                       coordinate_mapping="ball_to_cube_volume_preserving", radius_search_ignore_query_points=True)
    out_a = a(tf, tp, tp, 0.2, inp_importance=torch.from_numpy(imp).to(cuda))
    b.build(6, cuda)
    b.kernel.data.copy_(a.kernel.data)
    out_b = b(tf, tp, tp, 0.2, inp_importance=torch.from_numpy(imp).to(cuda))
    assert a.nns.neighbors_distance.numel() == a.nns.neighbors_index.numel()
    close(out_a, out_b.cpu().numpy())
    # explicit neighbour list + importance (utils/convolutions.py:341-349)
    nns = ops.fixed_radius_search(tp, tp, 0.1, ignore_query_point=True)
    w = win(nns.neighbors_distance / 0.1 ** 2)
    out_c = a(tf, tp, tp, 0.2, inp_importance=torch.from_numpy(imp).to(cuda), user_neighbors_index=nns.neighbors_index,
              user_neighbors_row_splits=nns.neighbors_row_splits, user_neighbors_importance=w)
    close(out_c, out_a.cpu().numpy())
    idx, splits, d2 = o64.fixed_radius_search(pts, pts, np.float32(0.1), True)
    ref = o64.continuous_conv(a.kernel.cpu().numpy(), pts, 0.2, None, pts, feats, imp, idx,
                              o64.window("poly6", d2 / np.float64(np.float32(0.1)) ** 2), splits,
                              coordinate_mapping="ball_to_cube_volume_preserving", normalize=False)
    close(out_a, ref)
    # a cell list built once can be handed in like the reference's fixed_radius_search_hash_table
    cl = ops.CellList(tp, 0.1)
    close(b(tf, tp, tp, 0.2, inp_importance=torch.from_numpy(imp).to(cuda), fixed_radius_search_hash_table=cl), ref)


def test_symmetric_circular_and_dense_center_options(cuda, cloud):
    from dmcf_b200.convolutions import ContinuousConv
    from dmcf_b200.losses import get_window_func
    rng, pts, feats, tp, tf = cloud
    # antisymmetric layer class == oracle two-pass form; stored kernel is halved on sym_axis
    s = ContinuousConv(3, [6, 6, 6], use_bias=False, symmetric=True, sym_axis=1, normalize=False,
                       coordinate_mapping="ball_to_cube_volume_preserving", window_function=get_window_func("peak"),
                       radius_search_ignore_query_points=True)
    out = s(tf, tp, tp, 0.2)
    assert tuple(s.kernel.shape) == (6, 3, 6, 6, 3)
    ref = o64.cconv_layer(feats, pts, pts, np.float32(0.2), s.kernel.cpu().numpy(), None, normalize=False,
                          coordinate_mapping="ball_to_cube_volume_preserving", ignore_query_points=True, window_name="peak",
                          symmetric=True, sym_axis=1)
    close(out, ref)
    assert float(out.double().sum(0).abs().max()) <= 1e-4 * float(out.double().abs().sum(0).max()) + 1e-6
    with pytest.raises(AssertionError):
        ContinuousConv(3, [5, 5, 5], symmetric=True, sym_axis=1).build(4, cuda)  # mirrored axis must be even
    # circular kernel: weights gathered by Chebyshev ring (utils/convolutions.py:395-407)
    c = ContinuousConv(2, [4, 4, 4], circular=True, normalize=False, use_bias=False)
    out = c(tf, tp, tp, 0.2)
    assert tuple(c.kernel.shape) == (2, 6, 2)
    k = c.kernel.cpu().numpy()
    g = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(4), indexing="ij"), -1)[..., ::-1] - 2.0 + 0.5
    ring = np.floor(np.abs(g)).max(-1).astype(int)
    ref = o64.cconv_layer(feats, pts, pts, np.float32(0.2), k[ring], None, normalize=False)
    close(out, ref)
    # dense layer for the centre point (utils/convolutions.py:462-464)
    d = ContinuousConv(4, [3, 3, 3], use_dense_layer_for_center=True, normalize=False)
    out = d(tf, tp, tp, 0.2)
    ref = o64.cconv_layer(feats, pts, pts, np.float32(0.2), d.kernel.cpu().numpy(), d.bias.cpu().numpy(), normalize=False)
    close(out, ref + feats.astype(np.float64) @ d.dense_kernel.cpu().numpy())
    # unreachable-from-DMCF options fail loudly
    with pytest.raises(NotImplementedError):
        ContinuousConv(2, [3, 3, 3], radius_search_metric="L1")
    with pytest.raises(NotImplementedError):
        d(tf, tp, tp, torch.full((700,), 0.2))


def test_point_sampling_layer(cuda, cloud):
    from dmcf_b200.convolutions import PointSampling
    from dmcf_b200.losses import get_window_func
    rng, pts, feats, tp, tf = cloud
    ps = PointSampling(window_function=get_window_func("poly6"), normalize=True)
    out = ps(tf, tp, tp[:100] + 0.01, 0.2)
    ref = o64.point_sampling(feats, pts, pts[:100] + np.float32(0.01), np.float32(0.2), window_name="poly6", normalize=True)
    close(out, ref)


@pytest.mark.parametrize("dim", [3, 2, 1])
@pytest.mark.parametrize("centralize", [True, False])
def test_grid_pos_matches_oracle_as_sets(cuda, dim, centralize):
    from dmcf_b200 import losses
    rng = np.random.default_rng(dim)
    pts = (rng.random((3000, 3)) * np.array([0.7, 0.5, 0.3])).astype(np.float32) - np.float32(0.2)
    vox = np.array([0.05, 0.05, 0.05], np.float32)
    if dim < 3:
        pts[:, 2] = 0; vox[2] = 0
    if dim < 2:
        pts[:, 0] = 0; vox[0] = 0
    t = torch.from_numpy(pts).to(cuda)
    for stride in (1.0, 2.0, 4.0):
        got = losses.grid_pos(t, vox * np.float32(stride), centralize=centralize).cpu().numpy()
        ref = o64.grid_pos(pts, vox * np.float32(stride), centralize=centralize)
        assert got.shape == ref.shape
        key = lambda a: a[np.lexsort((a[:, 0], a[:, 1], a[:, 2]))]
        assert np.array_equal(key(got), key(ref))  # bit-exact lattice positions, order aside
    dil, pcnt, idx = losses.get_dilated_pos(t, [1, 2, 4], voxel_size=[0.025 if v > 0 else 0.0 for v in vox], centralize=centralize)
    ref = o64.get_dilated_pos(pts, [1, 2, 4], voxel_size=[0.025 if v > 0 else 0.0 for v in vox], centralize=centralize)
    assert dil[0] is t and [d.shape[0] for d in dil] == [r.shape[0] for r in ref] == pcnt
    # voxel_size None: nested farthest-point subsets of N // stride points (utils/tools/losses.py:274-282)
    dil, pcnt, idx = losses.get_dilated_pos(t, [1, 2, 8], voxel_size=None)
    ref, ridx = o64.get_dilated_pos(pts, [1, 2, 8], voxel_size=None, return_idx=True)
    assert pcnt == [3000, 1500, 375] and idx[0] is None and tuple(idx[1].shape) == (1, 1500)
    for s in (1, 2):
        assert np.array_equal(idx[s][0].cpu().numpy(), ridx[s]) and np.array_equal(dil[s].cpu().numpy(), ref[s])


def test_compute_density_and_empty_inputs(cuda, cloud):
    from dmcf_b200 import losses, ops
    rng, pts, feats, tp, tf = cloud
    dens = losses.compute_density(tp, radius=0.1, win=losses.get_window_func("poly6")).cpu().numpy()
    idx, splits, d2 = o64.fixed_radius_search(pts, pts, np.float32(0.1))
    w = o64.window("poly6", d2 / np.float64(np.float32(0.1)) ** 2)
    ref = np.add.reduceat(np.concatenate([w, [0]]), splits[:-1]) * (np.diff(splits) > 0)
    close(dens, ref, 4.0)
    # zero out points / zero neighbours
    e = torch.zeros((0, 3), device=cuda)
    nns = ops.fixed_radius_search(tp, e, 0.1)
    out = ops.continuous_conv(torch.zeros((3, 3, 3, 6, 4), device=cuda), e, 0.2, None, tp, tf, None, nns.neighbors_index, None,
                              nns.neighbors_row_splits)
    assert tuple(out.shape) == (0, 4)
    far = tp[:5] + 100.0
    nns = ops.fixed_radius_search(tp, far, 0.1)
    out = ops.continuous_conv(torch.ones((3, 3, 3, 6, 4), device=cuda), far, 0.2, None, tp, tf, None, nns.neighbors_index, None,
                              nns.neighbors_row_splits, bias=torch.arange(4, device=cuda, dtype=torch.float32), normalize=True)
    assert torch.equal(out, torch.arange(4, device=cuda, dtype=torch.float32).expand(5, 4))


def test_canyon_inflow_rollout_growing_particle_set(cuda):
    """run_sample.py:121-181: the particle set grows by the inflow block every odd step while t < inflow."""
    import os
    from dmcf_b200 import config
    from dmcf_b200.simulator import Simulator
    sys_path = os.path.join(os.path.dirname(__file__), "golden")
    z = np.load(os.path.join(sys_path, "canyon_crop.npz"))
    import test_models_gpu as T
    model = config.build_model(T.liquid3d_cfg())
    sim = Simulator(model, device="cuda")
    model.load_weights(T.load_npz_weights("ckpt_Liquid3d.npz"), device=cuda)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)
    in_pos, in_vel = t(z["pos"]), t(z["vel"] + np.array([10.0, 0, -6.0], np.float32))
    sample = [in_pos, in_vel, None, None, t(z["box"]), t(z["box_normals"])]
    counts = []
    for step in range(5):
        sample = sim.step(sample)
        counts.append(sample[0].shape[0])
        assert torch.isfinite(sample[0]).all()
        if step % 2 == 1 and step < 4:
            sample[0] = torch.cat([sample[0], in_pos], dim=0)
            sample[1] = torch.cat([sample[1], in_vel], dim=0)
    assert counts == [1280, 1280, 2560, 2560, 3840]
    # the jet moves along +x / -z with the prescribed inflow velocity (10, 0, -6) * dt per step
    d = (sample[0][:1280].mean(0) - in_pos.mean(0)).cpu().numpy()
    assert d[0] > 0.5 and d[2] < -0.3


def test_run_sample_rollout_cli_body(cuda):
    """run_sample.run_rollout (the body of the reference's run_sample.py:140-181) on the canyon crop: per-particle
    acceleration tensor like the reference, inflow every odd step, results are the position sets per frame."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import run_sample
    from dmcf_b200 import config
    from dmcf_b200.simulator import Simulator
    import test_models_gpu as T
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "canyon_crop.npz"))
    model = config.build_model(T.liquid3d_cfg())
    sim = Simulator(model, device="cuda")
    model.load_weights(T.load_npz_weights("ckpt_Liquid3d.npz"), device=cuda)
    frame = dict(pos=z["pos"], vel=z["vel"], box=z["box"], box_normals=z["box_normals"])
    with torch.no_grad():
        res = run_sample.run_rollout(sim, model, frame, timesteps=6, inflow=4)
    assert [r.shape[0] for r in res] == [1280, 1280, 1280, 2560, 2560, 3840]
    assert all(torch.isfinite(r).all() for r in res)
    # first step equals Simulator.step with the same inputs (acc tensor == constant gravity vector)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)
    ref = sim.step([t(z["pos"]), t(z["vel"] + np.array([10.0, 0, -6.0], np.float32)), None, None, t(z["box"]), t(z["box_normals"])])
    assert torch.allclose(res[1], ref[0], rtol=0, atol=2e-6)


def test_density_loss_and_rollout_metrics(cuda):
    """utils/tools/losses.py:380-398 on the GPU vs the oracle's density; run_valid's metric dict is finite."""
    from dmcf_b200 import metrics
    from dmcf_b200.losses import get_window_func
    rng = np.random.default_rng(3)
    gt = (rng.random((700, 3)) * 0.3).astype(np.float32)
    pred = (gt + rng.normal(0, 0.004, gt.shape)).astype(np.float32)
    box = (rng.random((100, 3)) * 0.3).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(cuda)
    r = 0.05
    got = float(metrics.density_loss(t(gt), t(pred), torch.cat([t(pred), t(box)]), torch.cat([t(gt), t(box)]), radius=r,
                                     win=get_window_func("poly6")))
    dp = o64.compute_density(pred, np.concatenate([gt, box]), r, "poly6")
    dg = o64.compute_density(gt, np.concatenate([pred, box]), r, "poly6")
    ref = np.maximum(dp - dg.max() - 0.01, 0).mean()
    assert abs(got - ref) <= 1e-5 * max(abs(ref), 1.0)
    got_max = float(metrics.density_loss(t(gt), t(pred), radius=r, win=get_window_func("poly6"), use_max=True))
    d1, d2 = o64.compute_density(pred, pred, r, "poly6"), o64.compute_density(gt, gt, r, "poly6")
    assert abs(got_max - abs(d1.max() - d2.max()) / d2.max()) <= 1e-5
    vel = rng.standard_normal(gt.shape).astype(np.float32)
    m = metrics.rollout_metrics(t(pred), t(vel), t(gt), t(vel * 1.1), t(box), split="valid")
    assert set(m) >= {"mse_val", "chamfer_val", "dens_val", "chamfer_val_2", "vel_diff_val", "emd"} and all(np.isfinite(v) for v in m.values())
    from oracle import pointset
    clipped = np.clip(pred, box.min(0), box.max(0))  # run_valid clamps the prediction into the box (pipelines/simulator.py:218-219)
    want = pointset.emd_loss(gt, clipped)
    assert abs(m["emd"] - want) <= 5e-4 * want


def test_device_metrics_match_the_host_definitions(cuda):
    """utils/evaluation_helper.py:25-28, 43-72 on the device (chamfer through dmcf_nn_distance, histogram KL in torch ops) against
    the NumPy / SciPy statements of the same file."""
    from dmcf_b200 import metrics
    rng = np.random.default_rng(8)
    a = rng.standard_normal((4000, 3)).astype(np.float32)
    b = (rng.standard_normal((3500, 3)) * 1.2 + 0.1).astype(np.float32)
    t = lambda x: torch.from_numpy(x).to(cuda)
    got = metrics.chamfer_distance(t(a), t(b))
    assert isinstance(got, torch.Tensor) and got.is_cuda
    want = metrics.chamfer_distance(a, b)
    assert np.abs(got.cpu().numpy() - want).max() <= 2e-6 * max(want.max(), 1.0)
    va, vb = a, (a * 1.1 + rng.normal(0, 0.05, a.shape)).astype(np.float32)
    kl_gpu, kl_cpu = metrics.compare_dist(t(va), t(vb)), metrics.compare_dist(va.astype(np.float64), vb.astype(np.float64))
    assert abs(kl_gpu - kl_cpu) <= 1e-9 + 1e-7 * abs(kl_cpu)
    assert metrics.compare_dist(t(va), t(va)) < 1e-12
    v2 = rng.standard_normal((3000, 2)).astype(np.float32)  # 2-D velocities (WBC-SPH)
    assert abs(metrics.compare_dist(t(v2), t(v2 * 0.9)) - metrics.compare_dist(v2.astype(np.float64), (v2 * 0.9).astype(np.float64))) <= 1e-7
