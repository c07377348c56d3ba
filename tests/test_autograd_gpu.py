"""Gradients of the continuous convolution (dmcf_b200/autograd.py).  The conv is bilinear in (features,
filter), so a gradient is right iff the adjoint identity  <d_out, conv(direction)> == <gradient, direction>  holds for
every direction; the right-hand sides are evaluated with the float64 oracle's FORWARD conv only."""
import numpy as np
import pytest
import torch

from oracle import o64

pytestmark = pytest.mark.gpu

MAP = "ball_to_cube_volume_preserving"


def _scene(rng, n_in, n_out, same):
    pts = (rng.random((n_in, 3)) * 0.5).astype(np.float32)
    outp = pts[:n_out].copy() if same else (rng.random((n_out, 3)) * 0.5).astype(np.float32)
    return pts, outp


@pytest.mark.parametrize("ks,cin,cout,same,relu,window,ignore_q", [
    ((4, 4, 4), 8, 8, True, True, "poly6", False),
    ((4, 4, 4), 32, 16, False, False, "poly6", False),
    ((3, 3, 3), 5, 6, True, True, None, True),
    ((1, 8, 8), 16, 4, True, False, "peak", True),
], ids=["444-same-relu", "444-cross", "333-generic", "188-ignoreq"])
def test_conv_gradients_satisfy_the_adjoint_identities(cuda, ks, cin, cout, same, relu, window, ignore_q):
    from dmcf_b200 import autograd, ops
    rng = np.random.default_rng(hash((ks, cin, cout)) % 2 ** 31)
    n_in, n_out = 500, 500 if same else 350
    pts, outp = _scene(rng, n_in, n_out, same)
    if ks[0] == 1:
        pts[:, 2] = 0
        outp[:, 2] = 0
    extent = np.float32(0.2)
    feats = rng.standard_normal((n_in, cin)).astype(np.float32)
    filt = rng.uniform(-0.5, 0.5, ks + (cin, cout)).astype(np.float32)
    d_out = rng.standard_normal((n_out, cout)).astype(np.float32)
    scale = 0.7
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    nns = ops.fixed_radius_search(t(pts), t(outp), float(np.float32(0.5) * extent), ignore_query_point=ignore_q)
    W = t(filt).requires_grad_(True)
    F = t(feats).requires_grad_(True)
    kw = dict(align_corners=True, coordinate_mapping=MAP, interpolation="linear", window=window, relu_input=relu, feat_scale=scale)
    out = autograd.continuous_conv(W, t(outp), float(extent), t(pts), F, nns.neighbors_index, nns.neighbors_row_splits,
                                   drop_self=ignore_q, **kw)
    # same values as the non-differentiable op, and patches @ W reproduces them
    plain = ops.continuous_conv(t(filt), t(outp), float(extent), None, t(pts), t(feats), None, nns.neighbors_index, None,
                                nns.neighbors_row_splits, **kw)
    assert torch.equal(out.detach(), plain)
    patches = ops.conv_patches(ks, t(outp), float(extent), t(pts), t(feats), nns.neighbors_index, nns.neighbors_row_splits,
                               align_corners=True, coordinate_mapping=MAP, interpolation="linear", window=window,
                               relu_input=relu, feat_scale=scale)
    via = patches @ t(filt).reshape(-1, cout)
    assert float((via - plain).abs().max()) <= 2e-5 * float(plain.abs().max()) + 1e-6
    (out * t(d_out)).sum().backward()
    dW, dF = W.grad.double().cpu().numpy(), F.grad.double().cpu().numpy()

    idx, splits, d2 = o64.fixed_radius_search(pts, outp, np.float32(0.5) * extent, ignore_query_point=ignore_q)
    imp = None if window is None else o64.window(window, d2.astype(np.float64) / np.float64(np.float32(0.5) * extent) ** 2)
    g = lambda f: scale * (np.maximum(f, 0) if relu else f)

    def conv(w, f):
        return o64.continuous_conv(w, outp, extent, (0, 0, 0), pts, f, None, idx, imp, splits, align_corners=True,
                                   coordinate_mapping=MAP, normalize=False, interpolation="linear")

    ref_scale = np.abs(conv(filt.astype(np.float64), g(feats.astype(np.float64)))).max()
    for k in range(4):  # filter directions
        dw = rng.standard_normal(filt.shape)
        lhs, rhs = (dW * dw).sum(), (d_out * conv(dw, g(feats.astype(np.float64)))).sum()
        assert abs(lhs - rhs) <= 1e-4 * (np.abs(d_out).sum() * ref_scale / np.abs(filt).max()) / np.sqrt(d_out.size) + 1e-4, (k, lhs, rhs)
    mask = (feats > 0).astype(np.float64) if relu else np.ones_like(feats, np.float64)
    for k in range(4):  # feature directions (chain rule through relu / scale)
        df = rng.standard_normal(feats.shape)
        lhs, rhs = (dF * df).sum(), (d_out * conv(filt.astype(np.float64), scale * mask * df)).sum()
        assert abs(lhs - rhs) <= 1e-4 * np.abs(d_out).sum() * ref_scale / np.sqrt(d_out.size) + 1e-4, (k, lhs, rhs)


def test_layer_training_path_incl_antisymmetric_layer(cuda):
    """ContinuousConv with trainable parameters: value equals the inference path, gradients of the stored (half) kernel,
    the bias and the features satisfy the adjoint identities against the oracle's layer; a few SGD steps reduce a loss."""
    from dmcf_b200.convolutions import ContinuousConv
    from dmcf_b200.losses import get_window_func
    rng = np.random.default_rng(5)
    n, cin, cout = 450, 8, 3
    pts = (rng.random((n, 3)) * 0.5).astype(np.float32)
    feats = np.abs(rng.standard_normal((n, cin))).astype(np.float32)
    extent = np.float32(0.2)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    layer = ContinuousConv(filters=cout, kernel_size=[6, 6, 6], use_bias=False, normalize=False, align_corners=True, coordinate_mapping=MAP,
                           interpolation="linear", window_function=get_window_func("peak"),
                           radius_search_ignore_query_points=True, symmetric=True, sym_axis=1)
    torch.manual_seed(0)
    with torch.no_grad():
        ref_out = layer(t(feats), t(pts), t(pts), float(extent))  # builds the layer, fused inference path
    half = layer.kernel.detach().cpu().numpy().astype(np.float64)
    layer.kernel.requires_grad_(True)
    F = t(feats).requires_grad_(True)
    out = layer(F, t(pts), t(pts), float(extent))
    assert float((out.detach() - ref_out).abs().max()) <= 2e-5 * float(ref_out.abs().max()) + 1e-6
    d_out = rng.standard_normal((n, cout)).astype(np.float32)
    (out * t(d_out)).sum().backward()
    dK, dF = layer.kernel.grad.double().cpu().numpy(), F.grad.double().cpu().numpy()

    def layer_o64(k, f):
        return o64.cconv_layer(f, pts, pts, extent, k, None, align_corners=True, coordinate_mapping=MAP,
                               interpolation="linear", normalize=False, ignore_query_points=True, window_name="peak",
                               symmetric=True, sym_axis=1)

    base = np.abs(layer_o64(half, feats.astype(np.float64))).max()
    for k in range(3):
        dk = rng.standard_normal(half.shape)
        lhs, rhs = (dK * dk).sum(), (d_out * layer_o64(dk, feats.astype(np.float64))).sum()
        assert abs(lhs - rhs) <= 2e-4 * np.abs(d_out).sum() * base / np.abs(half).max() / np.sqrt(d_out.size) + 1e-4, (lhs, rhs)
        df = rng.standard_normal(feats.shape)
        lhs, rhs = (dF * df).sum(), (d_out * layer_o64(half, df)).sum()
        assert abs(lhs - rhs) <= 2e-4 * np.abs(d_out).sum() * base / np.sqrt(d_out.size) + 1e-4, (lhs, rhs)
    # a few optimisation steps on a regression target
    target = t(rng.standard_normal((n, cout)).astype(np.float32) * 0.01)
    opt = torch.optim.SGD([layer.kernel], lr=0.5)
    losses = []
    for _ in range(5):
        opt.zero_grad()
        layer._eff_cache = None
        loss = ((layer(t(feats), t(pts), t(pts), float(extent)) - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0]


def test_model_gradients_and_train_step(cuda):
    """Whole SymNet on the layer-by-layer path: d loss / d parameters from autograd vs central differences of the float64
    oracle model along random directions (a conv kernel of the stack, a Dense kernel, the antisymmetric half kernel, an
    input-conv bias), then a few optimisation steps of training.train_step reduce the loss."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import test_models_gpu as T
    from dmcf_b200 import config, scenes, training
    cfg = dict(name="SymNet", layer_channels=[[[4]], [[8]], [[8]], [[3]]], kernel_size=[4, 4, 4], sym_kernel_size=[6, 6, 6],
               coordinate_mapping=MAP, interpolation="linear", window="poly6", window_sym="peak", strides=[1],
               particle_radii=[0.1], timestep=0.02, grav=-9.81, out_scale=[0.0078125] * 3, sym_axis=1, add_merge=True,
               use_acc=False)
    scene = scenes.lattice_scene((6, 5, 5), dx=0.05, seed=9, open_top=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)
    model = config.build_model(cfg)
    model.init_weights(seed=4, device=cuda, scale=0.2)
    model.set_trainable(True)
    sample = [t(scene["pos"]), t(scene["vel"]), None, None, t(scene["box"]), t(scene["box_normals"])]
    rng = np.random.default_rng(1)
    target = scene["pos"] + np.array([0, -0.004, 0], np.float32) + rng.normal(0, 0.002, scene["pos"].shape).astype(np.float32)
    pos, vel = model(sample, training=True)
    loss = model.loss([pos, vel], [sample, t(target), sample[0], 0])["mse"]
    loss.backward()
    names = {n: l for n, l in model.named_layers().items()}
    weights = T.oracle_weights(model)

    def oracle_loss(w):
        p, _ = o64.ModelO64(cfg, w)(scene["pos"], scene["vel"], None, scene["box"], scene["box_normals"])
        return np.mean((((target.astype(np.float64) - p) ** 2).sum(-1) + 1e-9) ** 0.5)

    assert abs(float(loss.detach()) - oracle_loss(weights)) <= 1e-5 * oracle_loss(weights) + 1e-7
    picks = [("_all_convs/3", "kernel"), ("denses/0/0/0/0", "kernel"), ("sym_convs/0", "kernel"), ("fluid_convs", "bias")]
    for lname, attr in picks:
        # a layer is addressed by its checkpoint name or one of its aliases; the oracle's weight dict holds both
        name, layer = next((n, l) for n, l in names.items() if n == lname or lname in getattr(l, "_aliases", []))
        keys = [k + "/" + attr for k in [name] + list(getattr(layer, "_aliases", [])) if k + "/" + attr in weights]
        g = getattr(layer, attr).grad.double().cpu().numpy()
        base = weights[keys[0]].astype(np.float64)
        d = rng.standard_normal(base.shape)
        eps = 1e-3 * max(np.abs(base).max(), 1e-2)
        wp, wm = dict(weights), dict(weights)
        for k in keys:
            wp[k], wm[k] = base + eps * d, base - eps * d
        fd = (oracle_loss(wp) - oracle_loss(wm)) / (2 * eps)
        for k in keys:
            wp[k], wm[k] = base + 0.25 * eps * d, base - 0.25 * eps * d
        fd2 = (oracle_loss(wp) - oracle_loss(wm)) / (0.5 * eps)
        an = (g * d).sum()
        # the loss is piecewise smooth (relu): accept the spread of the two difference quotients as their uncertainty
        assert abs(an - fd2) <= 5e-3 * max(abs(fd2), abs(an)) + 2 * abs(fd - fd2) + 1e-8, (lname, an, fd, fd2)
    opt, sched = model.get_optimizer({"lr_boundaries": [1000], "lr_values": [1e-3, 1e-4]})
    tgt = torch.stack([sample[0], t(target)])
    losses = [training.train_step(model, opt, [sample], [tgt], grad_clip_norm=1.0, scheduler=sched)[0] for _ in range(6)]
    assert losses[-1] < losses[0]


def test_two_unrolled_steps_carry_the_gradient_through_pos_and_vel(cuda):
    """pipelines/simulator.py:316-421 unrolls T steps under one GradientTape: the loss of step 2 reaches the parameters through
    the step-1 outputs as well (pos2 = pos1 + dt * (vel1 + dt * g) + corr2, vel1 = (pos1 - pos0) / dt: d pos2 / d corr1 = 2 I plus
    the velocity features).  The integrate / correct kernels have no autograd, so on the training path they must run as torch
    ops.  Check: gradient of the 2-step loss == direct step-2 gradient + step-1 vector-Jacobian product with the cotangents
    autograd reports for the step-2 inputs; the position cotangent is exactly d loss / d pos2 (the conv geometry carries no
    position gradient, like Open3D's op)."""
    from dmcf_b200 import config, scenes
    cfg = dict(name="SymNet", layer_channels=[[[4]], [[8]], [[8]], [[3]]], kernel_size=[4, 4, 4], sym_kernel_size=[6, 6, 6],
               coordinate_mapping=MAP, interpolation="linear", window="poly6", window_sym="peak", strides=[1],
               particle_radii=[0.1], timestep=0.02, grav=-9.81, out_scale=[0.0078125] * 3, sym_axis=1, add_merge=True,
               use_acc=False)
    scene = scenes.lattice_scene((6, 5, 5), dx=0.05, seed=9, open_top=True)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)
    model = config.build_model(cfg)
    model.init_weights(seed=4, device=cuda, scale=0.2)
    model.set_trainable(True)
    params = [p for p in model.parameters() if p.requires_grad]
    sample = [t(scene["pos"]), t(scene["vel"]), None, None, t(scene["box"]), t(scene["box_normals"])]
    target = t(scene["pos"] + np.array([0, -0.012, 0], np.float32))

    def loss_of(pos):
        return ((((target - pos) ** 2).sum(-1) + 1e-9) ** 0.5).mean()

    # (a) two unrolled steps, one backward
    p1, v1 = model(sample, training=True)
    p2, _ = model([p1, v1] + sample[2:], training=True)
    g_total = torch.autograd.grad(loss_of(p2), params, allow_unused=True, retain_graph=True)  # step 1's graph is used again in (c)
    # (b) step 2 alone on leaf copies of the step-1 outputs
    p1l, v1l = p1.detach().requires_grad_(True), v1.detach().requires_grad_(True)
    p2b, _ = model([p1l, v1l] + sample[2:], training=True)
    p2b.retain_grad()
    lb = loss_of(p2b)
    g2 = torch.autograd.grad(lb, params + [p1l, v1l, p2b], allow_unused=True)
    c_pos, c_vel, dl_dp2 = g2[-3], g2[-2], g2[-1]
    assert torch.equal(c_pos, dl_dp2)  # identity path only
    assert float((c_vel - model.timestep * dl_dp2).abs().max()) > 0  # plus the velocity features of step 2
    # (c) step-1 vector-Jacobian product with those cotangents
    g1 = torch.autograd.grad([p1, v1], params, grad_outputs=[c_pos, c_vel], allow_unused=True)
    n_through = 0
    for gt, ga, gb in zip(g_total, g2[:len(params)], g1):
        if gt is None:
            assert ga is None and gb is None
            continue
        exp = (ga if ga is not None else 0) + (gb if gb is not None else 0)
        assert torch.allclose(gt, exp, rtol=1e-4, atol=1e-7 + 1e-5 * float(exp.abs().max()))
        if gb is not None and float(gb.abs().max()) > 0:
            n_through += 1
    assert n_through >= 4  # the step-1 path is really there (it was silently cut when integrate ran as a raw kernel)
