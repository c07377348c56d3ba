"""The reference's in-repo point-set ops on the C ABI, with the reference's Python signatures:

  * ``farthest_point_sample(npoint, inp)``, ``gather_point(inp, idx)``     <- utils/tools/sampling.py:62-112
  * ``approx_match(xyz1, xyz2, n=None, m=None)``, ``match_cost(...)``       <- utils/tools/tf_approxmatch.py:43-75
  * ``nn_distance(xyz1, xyz2)``                                             <- utils/tools/nn_distance.py:41-52
  * ``match_cost_grad`` (the registered gradient of ``match_cost``)         <- utils/tools/tf_approxmatch.py:78-88
  * ``emd_loss`` / ``approx_vel``                                           <- utils/tools/losses.py:401-413

Batched CUDA float32 tensors in the reference's layouts ([b, n, 3] point sets, match [b, m, n]); every function raises
on non-CUDA input -- there is no CPU path.  ``approx_match`` follows the reference's CUDA kernel (annealing levels
7..-2); ``first_level=8`` gives the schedule of its CPU kernel, which the tests pin against the reference's own
``approxmatch_cpu`` (oracle/_ref).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check
from .ops import _p, _req, _stream


def _points(t, name):
    _req(t, name, dim=3)
    if t.shape[2] != 3:
        raise ValueError(f"{name} must have shape [b, n, 3], got {tuple(t.shape)}")
    return t.contiguous()


def farthest_point_sample(npoint, inp, cluster_size=0):
    """utils/tools/sampling.py:86-94: inp [b, n, 3] -> int32 [b, npoint]; index 0 first, then the farthest point from
    the chosen set (float32 distances and tie order of utils/tools/sampling.cu:125-190)."""
    lib = _lib.load()
    inp = _points(inp, "inp")
    b, n = inp.shape[0], inp.shape[1]
    npoint = int(npoint)
    if n < 1 or npoint < 0 or npoint > n:
        raise ValueError(f"farthest_point_sample: need 0 <= npoint <= n and n >= 1 (n={n}, npoint={npoint})")
    idx = torch.empty((b, npoint), dtype=torch.int32, device=inp.device)
    if b == 0 or npoint == 0:
        return idx
    temp = torch.empty((b, n), dtype=torch.float32, device=inp.device)
    check(lib.dmcf_farthest_point_sample(_p(inp), b, n, npoint, _p(temp), _p(idx), int(cluster_size), _stream()))
    return idx


def gather_point(inp, idx):
    """utils/tools/sampling.py:62-70: inp [b, n, c], idx [b, m] -> [b, m, c] (plain indexing: torch's gather kernel)."""
    _req(inp, "inp", dim=3)
    _req(idx, "idx", dtype=torch.int32, dim=2)
    return torch.gather(inp, 1, idx.long().unsqueeze(-1).expand(-1, -1, inp.shape[2]))


def _counts(c, b, full, name):
    if c is None:
        return [full] * b
    c = [int(v) for v in (c.tolist() if isinstance(c, torch.Tensor) else c)]
    if len(c) != b or any(v < 1 or v > full for v in c):
        raise ValueError(f"{name}: per-item counts must be in [1, {full}] for each of the {b} batch items")
    return c


def approx_match(xyz1, xyz2, n=None, m=None, first_level=7):
    """utils/tools/tf_approxmatch.py:43-57: xyz1 [b, n, 3], xyz2 [b, m, 3] (optional per-item counts n, m [b]) ->
    match [b, m, n] (rows / columns beyond the counts stay zero)."""
    lib = _lib.load()
    xyz1, xyz2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    b, nn, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    if xyz2.shape[0] != b:
        raise ValueError("approx_match: batch sizes differ")
    cn, cm = _counts(n, b, nn, "n"), _counts(m, b, mm, "m")
    match = torch.zeros((b, mm, nn), dtype=torch.float32, device=xyz1.device)
    for i in range(b):
        ws_bytes = lib.dmcf_approx_match_workspace_bytes(cn[i], cm[i])
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz1.device)
        check(lib.dmcf_approx_match(_p(xyz1[i]), cn[i], _p(xyz2[i]), cm[i], int(first_level), _p(match[i]), nn, None, _p(ws),
                                    ws_bytes, _stream()))
    return match


def match_cost(xyz1, xyz2, match):
    """utils/tools/tf_approxmatch.py:63-72: -> cost [b] = sum_kl |xyz1_k - xyz2_l| match[l, k]."""
    lib = _lib.load()
    xyz1, xyz2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    _req(match, "match", dim=3)
    b, nn, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    if tuple(match.shape) != (b, mm, nn):
        raise ValueError(f"match must have shape {(b, mm, nn)}, got {tuple(match.shape)}")
    match = match.contiguous()
    cost = torch.empty(b, dtype=torch.float32, device=xyz1.device)
    ws_bytes = lib.dmcf_match_cost_workspace_bytes(nn, mm)
    ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=xyz1.device)
    for i in range(b):
        check(lib.dmcf_match_cost(_p(xyz1[i]), nn, _p(xyz2[i]), mm, _p(match[i]), nn, _p(cost[i:i + 1]), _p(ws), ws_bytes,
                                  _stream()))
    return cost


def match_cost_grad(xyz1, xyz2, match):
    """op MatchCostGrad (utils/tools/tf_approxmatch.py:78-88 registers it as the gradient of match_cost): d cost / d xyz1
    [b, n, 3] and d cost / d xyz2 [b, m, 3] with the match held constant."""
    lib = _lib.load()
    xyz1, xyz2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    _req(match, "match", dim=3)
    b, nn, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    if tuple(match.shape) != (b, mm, nn):
        raise ValueError(f"match must have shape {(b, mm, nn)}, got {tuple(match.shape)}")
    match = match.contiguous()
    g1, g2 = torch.empty_like(xyz1), torch.empty_like(xyz2)
    for i in range(b):
        check(lib.dmcf_match_cost_grad(_p(xyz1[i]), nn, _p(xyz2[i]), mm, _p(match[i]), nn, _p(g1[i]), _p(g2[i]), _stream()))
    return g1, g2


class _MatchCost(torch.autograd.Function):
    """match_cost with the reference's registered gradient (the match itself carries none: ops.NoGradient('ApproxMatch'))."""

    @staticmethod
    def forward(ctx, xyz1, xyz2, match):
        ctx.save_for_backward(xyz1, xyz2, match)
        return match_cost(xyz1, xyz2, match)

    @staticmethod
    def backward(ctx, grad_cost):
        xyz1, xyz2, match = ctx.saved_tensors
        g1, g2 = match_cost_grad(xyz1.detach(), xyz2.detach(), match)
        gc = grad_cost.reshape(-1, 1, 1)
        return g1 * gc, g2 * gc, None


def emd_cost(xyz1, xyz2, n=None, m=None, first_level=7):
    """``match_cost(xyz1, xyz2, approx_match(xyz1, xyz2, n, m))`` without the [m, n] match matrix: the cost is accumulated
    while the assignment is annealed (O(n + m) memory; the metric of run_valid on scenes whose match matrix would not fit)."""
    lib = _lib.load()
    xyz1, xyz2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    b, nn, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    if xyz2.shape[0] != b:
        raise ValueError("emd_cost: batch sizes differ")
    cn, cm = _counts(n, b, nn, "n"), _counts(m, b, mm, "m")
    cost = torch.empty(b, dtype=torch.float32, device=xyz1.device)
    for i in range(b):
        ws_bytes = lib.dmcf_approx_match_workspace_bytes(cn[i], cm[i])
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=xyz1.device)
        check(lib.dmcf_approx_match(_p(xyz1[i]), cn[i], _p(xyz2[i]), cm[i], int(first_level), None, 0, _p(cost[i:i + 1]), _p(ws),
                                    ws_bytes, _stream()))
    return cost


def emd_loss(y_true, y_pred, n=None, m=None, fused=True, first_level=7):
    """utils/tools/losses.py:401-408: match cost / max(n, m), [b]."""
    b = y_true.shape[0]
    cn, cm = _counts(n, b, y_true.shape[1], "n"), _counts(m, b, y_pred.shape[1], "m")
    if y_true.requires_grad or y_pred.requires_grad:  # training loss (utils/tools/losses.py:105-106)
        with torch.no_grad():
            match = approx_match(y_true, y_pred, n, m, first_level)
        cost = _MatchCost.apply(y_true, y_pred, match)
    elif fused:
        cost = emd_cost(y_true, y_pred, n, m, first_level)
    else:
        cost = match_cost(y_true, y_pred, approx_match(y_true, y_pred, n, m, first_level))
    denom = torch.tensor([float(max(a, c)) for a, c in zip(cn, cm)], dtype=torch.float32, device=cost.device)
    return cost / denom


def approx_vel(pos_0, pos_1, n=None, m=None):
    """utils/tools/losses.py:411-414: sum_l match[l, k] (pos_1[l] - pos_0[k]) -> [b, n, 3]."""
    match = approx_match(pos_0, pos_1, n, m)  # [b, m, n]
    return torch.bmm(match.transpose(1, 2), pos_1) - pos_0 * match.sum(dim=1).unsqueeze(-1)


def nearest_distance(queries, points):
    """For every row of ``queries`` [n, 3] the squared distance to / index of its nearest row of ``points`` [m, 3] (one
    direction of NnDistance; the chamfer metric of utils/evaluation_helper.py:25-28 on the device)."""
    lib = _lib.load()
    q, p = _points(queries.unsqueeze(0), "queries")[0], _points(points.unsqueeze(0), "points")[0]
    n, m = q.shape[0], p.shape[0]
    d = torch.empty(n, dtype=torch.float32, device=q.device)
    i = torch.empty(n, dtype=torch.int32, device=q.device)
    if n and not m:
        raise ValueError("nearest_distance: the reference point set is empty")
    if n:
        check(lib.dmcf_nn_distance(_p(q), n, _p(p), m, _p(d), _p(i), _stream()))
    return d, i


def nn_distance(xyz1, xyz2):
    """utils/tools/nn_distance.py:41-52: (dist1 [b, n], idx1 [b, n], dist2 [b, m], idx2 [b, m]); squared distances."""
    lib = _lib.load()
    xyz1, xyz2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    b, nn, mm = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    if xyz2.shape[0] != b:
        raise ValueError("nn_distance: batch sizes differ")
    if (nn == 0) != (mm == 0):
        raise ValueError("nn_distance: one point set is empty")
    dev = xyz1.device
    d1, i1 = torch.empty((b, nn), dtype=torch.float32, device=dev), torch.empty((b, nn), dtype=torch.int32, device=dev)
    d2, i2 = torch.empty((b, mm), dtype=torch.float32, device=dev), torch.empty((b, mm), dtype=torch.int32, device=dev)
    if nn == 0:
        return d1, i1, d2, i2
    for i in range(b):
        check(lib.dmcf_nn_distance(_p(xyz1[i]), nn, _p(xyz2[i]), mm, _p(d1[i]), _p(i1[i]), _stream()))
        check(lib.dmcf_nn_distance(_p(xyz2[i]), mm, _p(xyz1[i]), nn, _p(d2[i]), _p(i2[i]), _stream()))
    return d1, i1, d2, i2
