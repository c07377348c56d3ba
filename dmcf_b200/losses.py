"""Host-side mirror of the hot-path pieces of the reference's ``utils/tools/losses.py``:
``get_window_func`` (:8-44), ``grid_pos`` (:136-181), ``get_dilated_pos`` (:249-284), ``compute_density`` (:287-308).
Training losses / EMD / chamfer of that file are out of scope."""
from __future__ import annotations

import math

import torch

from . import ops

_TYPES = ("poly6", "cubic", "linear", "peak", "cubic_grad")


class WindowFunction:
    """Callable radial window on q = d^2/r^2 (torch tensors).  Carries ``typ``/``fac`` so ContinuousConv can
    evaluate it inside the CUDA kernel instead of materialising a per-pair importance array."""

    def __init__(self, typ, fac=1.0):
        if typ not in _TYPES:
            raise NotImplementedError(typ)
        self.typ = typ
        self.fac = float(fac)

    def __call__(self, q):
        fac, typ = self.fac, self.typ
        if typ == "poly6":
            return fac * torch.clamp((1 - q) ** 3, 0, 1)
        qs = torch.sqrt(q)
        if typ == "cubic":
            return fac * 4 / 3 * torch.where(q <= 1, torch.where(qs <= 0.5, 6 * (qs ** 3 - q) + 1, 2 * (1 - qs) ** 3),
                                             torch.zeros_like(qs))
        if typ == "linear":
            return fac * (1 - qs)
        if typ == "peak":
            return fac * (1 - 2 * qs + q)
        return fac * 4 / 3 * torch.where(q <= 1, torch.where(qs <= 0.5, 18 * q - 12 * qs, -6 * (1 - qs) ** 2),
                                         torch.zeros_like(qs))


def get_window_func(typ, fac=1.0, **kwargs):
    """utils/tools/losses.py:8-44: returns None for ``typ is None``."""
    if typ is None:
        return None
    return WindowFunction(typ, fac)


def point_mean(pos):
    """Mean position used by ``centralize`` (utils/tools/losses.py:137-139).  Accumulated in float64 and rounded
    once so that every implementation in this repo (CUDA, oracles, slab ranks) gets bit-identical lattices."""
    m = ops.valid_rows_mask(pos)
    if m is None:
        return pos.to(torch.float64).mean(dim=0).to(torch.float32)
    # capacity-sized buffer: mean over the valid rows only (count on the device, no host sync)
    p64 = torch.where(m[:, None], pos.to(torch.float64), torch.zeros((), dtype=torch.float64, device=pos.device))
    return (p64.sum(dim=0) / ops.count_of(pos).to(torch.float64)).to(torch.float32)


def grid_pos(pos, voxel_size, centralize=False, pad=0, hyst=0.1, center=None):
    """Lattice points (cell pitch ``voxel_size``) touched by any particle, utils/tools/losses.py:136-181.
    Output order is ascending linear voxel id (the reference's first-occurrence order is an internal detail,
    SURVEY A.6)."""
    if pad != 0:
        raise NotImplementedError("sample_pad != 0 is not used by any shipped config")
    v = [float(x) for x in torch.as_tensor(voxel_size, dtype=torch.float32).reshape(3).tolist()]
    if centralize and center is None:
        center = point_mean(pos)
    return ops.grid_pos(pos, v, center if centralize else None, float(hyst))


def get_dilated_pos(pos, strides, voxel_size=None, centralize=False, pad=0, hyst=0.1):
    """utils/tools/losses.py:249-284.  Returns (positions per scale, counts per scale, idx) like the reference: voxel
    lattices (``grid_pos``) when ``voxel_size`` is given, else nested farthest-point subsets with ``idx[s]`` = int32
    [1, N_s] indices of scale s inside scale s-1 (``idx`` is only filled in that mode, like the reference)."""
    dilated, pcnt, idx = [], [], []
    center = None
    for stride in strides:
        if stride == 1:
            dilated.append(pos)
            pcnt.append(pos.shape[0])
            idx.append(None)
        else:
            if voxel_size is None:  # :274-282: N // stride farthest points of the PREVIOUS scale (nested subsets)
                from .pointops import farthest_point_sample, gather_point
                sample_cnt = max(pos.shape[0] // int(stride), 1)
                pcnt.append(sample_cnt)
                idx.append(farthest_point_sample(sample_cnt, dilated[-1].unsqueeze(0)))
                dilated.append(gather_point(dilated[-1].unsqueeze(0), idx[-1])[0])
                continue
            if centralize and center is None:
                center = point_mean(pos)
            vs = torch.as_tensor(voxel_size, dtype=torch.float32).reshape(3) * float(stride)
            dilated.append(grid_pos(pos, vs, centralize, pad, hyst, center))
            pcnt.append(dilated[-1].shape[0])
    return dilated, pcnt, idx


def compute_density(out_pos, in_pos=None, radius=0.005, win=None):
    """utils/tools/losses.py:287-308: sum over neighbours of win(d^2/r^2)."""
    if in_pos is None:
        in_pos = out_pos
    if win is None:
        win = lambda x: x  # noqa: E731  (the reference warns and uses the identity)
    nns = ops.fixed_radius_search(in_pos, out_pos, radius, return_distances=True)
    r = torch.tensor(float(radius), dtype=torch.float32, device=out_pos.device)
    w = win(nns.neighbors_distance / (r * r))
    csum = torch.zeros(w.shape[0] + 1, dtype=torch.float64, device=out_pos.device)
    csum[1:] = torch.cumsum(w.to(torch.float64), 0)
    rs = nns.neighbors_row_splits
    return (csum[rs[1:]] - csum[rs[:-1]]).to(torch.float32)


def get_loss(typ, fac=1.0, **kwargs):
    """Training losses of utils/tools/losses.py:47-134 that need nothing but the predicted / target positions:
    'mse', 'weighted_mse', 'vel', 'dense' (density_loss) and 'emd' (approx-match cost with the reference's gradient)."""
    gamma = kwargs.get("gamma", 0.5)

    def pre_factor(kw):
        return math.exp(-kwargs.get("pre_scale", 0.0) * float(kw.get("pre_steps") or 0))

    if typ == "mse":
        return lambda target, pred, **kw: fac * (pre_factor(kw) * (((target - pred) ** 2).sum(-1) + 1e-9) ** gamma).mean()
    if typ == "weighted_mse":
        def f(target, pred, **kw):
            importance = torch.exp(-kwargs.get("neighbor_scale", 1.0) * kw.get("num_fluid_neighbors"))
            return fac * (pre_factor(kw) * importance * (((target - pred) ** 2).sum(-1) + 1e-9) ** gamma).mean()
        return f
    if typ == "vel":
        def f(target, pred, **kw):
            inp, prev = kw.get("input")[0], kw.get("target_prev")
            return fac * ((((target - prev) - (pred - inp)) ** 2).sum(-1) + 1e-9).pow(gamma).mean()
        return f
    if typ == "dense":
        from functools import partial
        from .metrics import density_loss
        kw2 = dict(kwargs)
        return partial(density_loss, win=get_window_func(kw2.pop("win", None)), **kw2)
    if typ == "emd":  # :105-106 -> emd_loss(y_true, y_pred) on [n, 3] frames
        from .pointops import emd_loss

        def f(target, pred, **kw):
            return fac * emd_loss(target.unsqueeze(0).contiguous(), pred.unsqueeze(0).contiguous()).mean()
        return f
    raise NotImplementedError(f"loss {typ!r}")
