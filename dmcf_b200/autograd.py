"""Gradients of the plain continuous convolution, so that the layer API can be trained with
``torch.autograd`` (the reference trains through Open3D's ``continuous_conv_transpose`` /
``continuous_conv_backprop_filter`` ops, cf. utils/convolutions.py:844-874; pipelines/simulator.py:316-421).

The conv is bilinear in (features, filter):  out = B(f) @ W  with the patch matrix B (dmcf_cconv_patches), hence
  d/dW  = B(f)^T @ d_out                      -- phase 1 of the forward kernel, then ONE plain GEMM per chunk of points;
  d/df  = the forward kernel itself on the TRANSPOSED neighbour list (roles of the two point sets swapped) with the
          filter mirrored along all three axes and transposed in (cin, cout): the coordinate mappings are odd and the
          interpolation grid is symmetric, so the weight of cell c for the offset r is the weight of the mirrored cell
          for -r; then the chain rule through relu / feat_scale.
Supported: window evaluated in the kernel (or none), relu_input, feat_scale, skip_self; not: normalize, per-point
importances, non-zero offset (they break the bilinear / mirrored form) -- DMCF's convs use none of them."""
from __future__ import annotations

import torch

from . import ops

PATCH_CHUNK_BYTES = 1 << 30  # patch rows materialised per GEMM (the reference batches its patch matrix the same way)


class ContinuousConvFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, filters, inp_features, out_positions, inp_positions, neighbors_index, neighbors_row_splits, extent,
                drop_self, kw):
        out = ops.continuous_conv(filters, out_positions, extent, None, inp_positions, inp_features, None, neighbors_index,
                                  None, neighbors_row_splits, **kw)
        ctx.save_for_backward(filters, inp_features, out_positions, inp_positions, neighbors_index, neighbors_row_splits)
        ctx.extent, ctx.kw, ctx.drop_self = float(extent), dict(kw), bool(drop_self)
        return out

    @staticmethod
    def backward(ctx, d_out):
        filters, feats, out_pos, inp_pos, nbr_index, row_splits = ctx.saved_tensors
        kw, extent = ctx.kw, ctx.extent
        d_out = d_out.contiguous()
        geo = dict(align_corners=kw.get("align_corners", True), coordinate_mapping=kw.get("coordinate_mapping", "ball_to_cube_radial"),
                   interpolation=kw.get("interpolation", "linear"), window=kw.get("window"), window_fac=kw.get("window_fac", 1.0),
                   skip_self=kw.get("skip_self", False))
        relu, scale = bool(kw.get("relu_input", False)), float(kw.get("feat_scale", 1.0))
        kz, ky, kx, cin, cout = filters.shape
        d_filters = d_feats = None
        if ctx.needs_input_grad[0]:
            kc = kz * ky * kx * cin
            n_out = out_pos.shape[0]
            chunk = max(1, min(n_out, PATCH_CHUNK_BYTES // (4 * kc)))
            acc = torch.zeros((kc, cout), dtype=torch.float32, device=filters.device)
            for a in range(0, n_out, chunk):
                b = min(n_out, a + chunk)
                patches = ops.conv_patches((kz, ky, kx), out_pos[a:b], extent, inp_pos, feats, nbr_index, row_splits[a:b + 1],
                                           relu_input=relu, feat_scale=scale, **geo)
                acc.addmm_(patches.t(), d_out[a:b])
            d_filters = acc.reshape(kz, ky, kx, cin, cout)
        if ctx.needs_input_grad[1]:
            # transposed list: for every INPUT point the out points that see it (same radius, same self-exclusion)
            nns_t = ops.fixed_radius_search(out_pos, inp_pos, 0.5 * extent, ignore_query_point=ctx.drop_self,
                                            return_distances=False)
            w_t = filters.flip(0, 1, 2).transpose(3, 4).contiguous()
            g = ops.continuous_conv(w_t, inp_pos, extent, None, out_pos, d_out, None, nns_t.neighbors_index, None,
                                    nns_t.neighbors_row_splits, normalize=False, **geo)
            if scale != 1.0:
                g = g * scale
            if relu:
                g = g * (feats > 0).to(g.dtype)
            d_feats = g
        return d_filters, d_feats, None, None, None, None, None, None, None


def continuous_conv(filters, out_positions, extent, inp_positions, inp_features, neighbors_index, neighbors_row_splits,
                    drop_self=False, **kw):
    """Differentiable ``ops.continuous_conv`` (w.r.t. ``filters`` and ``inp_features``).  ``drop_self``: the neighbour
    list was built with ignore_query_point (the transposed list of the backward pass is then built the same way)."""
    unsupported = [k for k in ("normalize", "ascc", "dense_cin", "accumulate") if kw.get(k)]
    unsupported += [k for k in ("bias", "residual", "dense_inp", "out", "nbr_range", "pair_records") if kw.get(k) is not None]
    if unsupported:
        raise NotImplementedError(f"continuous_conv gradients do not support {unsupported}")
    return ContinuousConvFunction.apply(filters, inp_features, out_positions, inp_positions, neighbors_index,
                                        neighbors_row_splits, float(extent), bool(drop_self), kw)
