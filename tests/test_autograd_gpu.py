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
