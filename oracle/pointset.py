"""CPU restatement of the reference's in-repo point-set ops and the ctypes front end of the
reference's OWN CPU functions compiled into ``oracle/_ref/libdmcf_refops.so`` (recipe: oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and bench.py's CPU legs, never by the product.

Pinning status of this file (unlike o64/o32, whose Open3D boundary stays PARITY UNPINNED):
* ``approx_match`` / ``match_cost`` / ``nn_search`` are checked against the reference's own code -- ``approxmatch_cpu``,
  ``matchcost_cpu`` (utils/tools/tf_approxmatch.cpp:51-112, 177-196) and ``nnsearch`` (utils/tools/nn_distance.cpp:47-70)
  compiled unmodified from the reference checkout (tests/test_pointset_cpu.py).
* ``farthest_point_sample`` restates the CUDA kernel utils/tools/sampling.cu:125-190 (the reference has no CPU twin):
  pinned only by the definition (every chosen point maximises the distance to the chosen set) -- "parity unpinned" for the
  tie order and the FMA contraction, which follow the kernel text and what nvcc 12.9 does with its distance expression.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

F32 = np.float32
_HERE = os.path.dirname(os.path.abspath(__file__))
REF_PATH = os.path.join(_HERE, "_ref", "libdmcf_refops.so")
_REF = None


# ---------------------------------------------------------------------------------------------------------
# the reference's own CPU functions (oracle/_ref)
# ---------------------------------------------------------------------------------------------------------
def ref_available():
    return os.path.exists(REF_PATH)


def ref_lib():
    global _REF
    if _REF is None:
        if not ref_available():
            raise FileNotFoundError(f"{REF_PATH} missing: `make -C oracle` builds it where /root/reference exists")
        L = C.CDLL(REF_PATH)
        vp, ci = C.c_void_p, C.c_int
        L.ref_approxmatch.argtypes = [ci, ci, ci, vp, vp, vp]
        L.ref_approxmatch_dyn.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp]
        L.ref_matchcost.argtypes = [ci, ci, ci, vp, vp, vp, vp]
        L.ref_matchcostgrad.argtypes = [ci, ci, ci, vp, vp, vp, vp, vp]
        L.ref_nnsearch.argtypes = [ci, ci, ci, vp, vp, vp, vp]
        for f in (L.ref_approxmatch, L.ref_approxmatch_dyn, L.ref_matchcost, L.ref_matchcostgrad, L.ref_nnsearch):
            f.restype = None
        _REF = L
    return _REF


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def ref_approx_match(xyz1, xyz2):
    """approxmatch_cpu (tf_approxmatch.cpp:51-112): xyz1 [b,n,3], xyz2 [b,m,3] -> match [b,m,n] (levels 8..-2, double)."""
    xyz1, xyz2 = np.ascontiguousarray(xyz1, F32), np.ascontiguousarray(xyz2, F32)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    match = np.zeros((b, m, n), F32)
    ref_lib().ref_approxmatch(b, n, m, _p(xyz1), _p(xyz2), _p(match))
    return match


def ref_approx_match_dyn(xyz1, xyz2, cn, cm):
    """approxmatch_cpu_dyn (:113-176) with per-item point counts."""
    xyz1, xyz2 = np.ascontiguousarray(xyz1, F32), np.ascontiguousarray(xyz2, F32)
    cn, cm = np.ascontiguousarray(cn, np.int32), np.ascontiguousarray(cm, np.int32)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    match = np.zeros((b, m, n), F32)
    ref_lib().ref_approxmatch_dyn(b, n, m, _p(xyz1), _p(xyz2), _p(match), _p(cn), _p(cm))
    return match


def ref_match_cost(xyz1, xyz2, match):
    """matchcost_cpu (:177-196) -> cost [b]."""
    xyz1, xyz2, match = (np.ascontiguousarray(a, F32) for a in (xyz1, xyz2, match))
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    cost = np.zeros(b, F32)
    ref_lib().ref_matchcost(b, n, m, _p(xyz1), _p(xyz2), _p(match), _p(cost))
    return cost


def ref_match_cost_grad(xyz1, xyz2, match):
    """matchcostgrad_cpu (:198-232) -> (grad1 [b,n,3], grad2 [b,m,3]).  (The reference zeroes only component 0 of grad1
    before accumulating; the buffers handed over here are zero-filled.)"""
    xyz1, xyz2, match = (np.ascontiguousarray(a, F32) for a in (xyz1, xyz2, match))
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    g1, g2 = np.zeros((b, n, 3), F32), np.zeros((b, m, 3), F32)
    ref_lib().ref_matchcostgrad(b, n, m, _p(xyz1), _p(xyz2), _p(match), _p(g1), _p(g2))
    return g1, g2


def ref_nn_search(xyz1, xyz2):
    """nnsearch (nn_distance.cpp:47-70): xyz1 [b,n,3], xyz2 [b,m,3] -> (squared distance [b,n], index [b,n])."""
    xyz1, xyz2 = np.ascontiguousarray(xyz1, F32), np.ascontiguousarray(xyz2, F32)
    b, n, m = xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]
    dist, idx = np.zeros((b, n), F32), np.zeros((b, n), np.int32)
    ref_lib().ref_nnsearch(b, n, m, _p(xyz1), _p(xyz2), _p(dist), _p(idx))
    return dist, idx


# ---------------------------------------------------------------------------------------------------------
# restatements
# ---------------------------------------------------------------------------------------------------------
def _fma32(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in the 64-bit mantissa of x87 long double."""
    ld = np.longdouble
    return (a.astype(ld) * b.astype(ld) + c.astype(ld)).astype(F32)


def fps_dist2(p, q):
    """Squared distance of the FPS / approx-match kernels as nvcc contracts (x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1)
    (sampling.cu:156, tf_approxmatch.cu:78): fma(dz,dz, fma(dx,dx, dy*dy)).  p: [..,3] points x2, q: [3] point x1."""
    d = (p - q).astype(F32)
    dx, dy, dz = d[..., 0], d[..., 1], d[..., 2]
    return _fma32(dz, dz, _fma32(dx, dx, (dy * dy).astype(F32)))


def farthest_point_sample(npoint, points, ref_threads=512):
    """utils/tools/sampling.cu:125-190 for one batch item: points [n,3] -> idx [npoint] int32.  Index 0 first; then the
    point with the largest running minimum distance; ties resolved like the kernel's strided scan + pairwise tree
    (smaller k mod blockDim first, then smaller k; blockDim = 512, sampling.cu:209)."""
    pts = np.ascontiguousarray(points, F32)
    n = pts.shape[0]
    idx = np.zeros(npoint, np.int32)
    if npoint == 0:
        return idx
    temp = np.full(n, 1e38, F32)
    ks = np.arange(n)
    old = 0
    for j in range(1, npoint):
        temp = np.minimum(fps_dist2(pts, pts[old]), temp)
        cand = ks[temp == temp.max()]
        old = int(cand[np.lexsort((cand, cand % ref_threads))[0]])
        idx[j] = old
    return idx


def approx_match(xyz1, xyz2, first_level=7, dtype=np.float64):
    """The soft assignment of tf_approxmatch.cu:27-160 for one batch item (xyz1 [n,3], xyz2 [m,3] -> match [m,n]),
    in the CUDA kernel's formulation (levels first_level..-2; the CPU twin tf_approxmatch.cpp:51-112 starts at 8 and
    places the 1e-9 guards slightly differently), evaluated in ``dtype``."""
    x1, x2 = np.asarray(xyz1, dtype), np.asarray(xyz2, dtype)
    n, m = x1.shape[0], x2.shape[0]
    multi_l, multi_r = (1.0, float(n // m)) if n >= m else (float(m // n), 1.0)
    remain_l, remain_r = np.full(n, multi_l, dtype), np.full(m, multi_r, dtype)
    d2 = ((x2[:, None, :] - x1[None, :, :]) ** 2).sum(-1)  # [m, n]
    match = np.zeros((m, n), dtype)
    eps = dtype(1e-9)
    for j in range(first_level, -3, -1):
        level = dtype(0.0) if j == -2 else dtype(-(4.0 ** j))
        e = np.exp(level * d2)
        suml = eps + (e * remain_r[:, None]).sum(0)
        ratio_l = remain_l / suml
        sumr = (e * ratio_l[None, :]).sum(1) * remain_r
        consumption = np.minimum(remain_r / (sumr + eps), dtype(1.0))
        ratio_r = consumption * remain_r
        remain_r = np.maximum(dtype(0.0), remain_r - sumr)
        w = e * ratio_l[None, :] * ratio_r[:, None]
        match += w
        remain_l = np.maximum(dtype(0.0), remain_l - w.sum(0))
    return match


def match_cost(xyz1, xyz2, match):
    """tf_approxmatch.cpp:177-196 for one batch item: sum_kl |x1_k - x2_l| match[l,k] (float64)."""
    x1, x2 = np.asarray(xyz1, np.float64), np.asarray(xyz2, np.float64)
    d = np.sqrt(((x2[:, None, :] - x1[None, :, :]) ** 2).sum(-1))
    return float((d * np.asarray(match, np.float64)).sum())


def match_cost_grad(xyz1, xyz2, match):
    """tf_approxmatch.cpp:198-232 for one batch item in float64: (grad1 [n,3], grad2 [m,3])."""
    x1, x2, mt = np.asarray(xyz1, np.float64), np.asarray(xyz2, np.float64), np.asarray(match, np.float64)
    diff = x2[:, None, :] - x1[None, :, :]  # [m, n, 3]
    d = np.maximum(np.sqrt((diff ** 2).sum(-1)), 1e-20)
    g = mt[:, :, None] * diff / d[:, :, None]
    return -g.sum(0), g.sum(1)


def emd_loss(y_true, y_pred, first_level=7):
    """utils/tools/losses.py:401-408 for one batch item: match cost / max(n, m)."""
    mt = approx_match(y_true, y_pred, first_level)
    return match_cost(y_true, y_pred, mt) / max(len(y_true), len(y_pred))


def nn_search(xyz1, xyz2):
    """utils/tools/nn_distance.cpp:47-70 for one batch item: float32 (x*x + y*y) + z*z, first minimum."""
    a, b = np.ascontiguousarray(xyz1, F32), np.ascontiguousarray(xyz2, F32)
    d = (b[None, :, :] - a[:, None, :]).astype(F32)
    d2 = ((d[..., 0] * d[..., 0]).astype(F32) + (d[..., 1] * d[..., 1]).astype(F32)).astype(F32)
    d2 = (d2 + (d[..., 2] * d[..., 2]).astype(F32)).astype(F32)
    idx = d2.argmin(1).astype(np.int32)  # first minimum
    return d2[np.arange(len(a)), idx], idx
