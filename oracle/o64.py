"""O64 -- NumPy restatement of DMCF's per-step particle hot path (TEST INFRASTRUCTURE ONLY).

This module is the *checker*, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under ``dmcf_b200/``
imports ``oracle``.

PARITY UNPINNED: the arithmetic of the path lives in the pip wheel ``open3d==0.15.2``
(``/root/reference/requirements.txt:2``) which is neither vendored in the reference nor installable in this
image (no TensorFlow 2.5 / Open3D wheels for Python 3.12, no network).  The Open3D parts below restate the
published algorithm of ``open3d.ml`` 0.15 (``FixedRadiusSearch``, ``ContinuousConv``: coordinate mappings,
trilinear interpolation, neighbour importance, normalisation).  Everything that *is* visible in the
reference is followed line by line and cited (paths relative to ``/root/reference``).

All functions take float32 inputs and compute in float64 unless stated, so that O64 can calibrate the
float32 implementations (O32 C oracle, CUDA kernels).
"""
from __future__ import annotations

import numpy as np
from scipy import sparse
from scipy.spatial import cKDTree

F32 = np.float32


# ----------------------------------------------------------------------------------------------------------
# window functions -- utils/tools/losses.py:8-44
# ----------------------------------------------------------------------------------------------------------
def window(name, q, fac=1.0):
    """Radial window on q = d^2 / r^2 (utils/tools/losses.py:8-44)."""
    if name is None:
        return None
    q = np.asarray(q, dtype=np.float64)
    if name == "poly6":  # :9-12
        return fac * np.clip((1 - q) ** 3, 0, 1)
    qs = np.sqrt(q)
    if name == "cubic":  # :13-20
        return fac * 4 / 3 * np.where(q <= 1, np.where(qs <= 0.5, 6 * (qs ** 3 - q) + 1, 2 * (1 - qs) ** 3), 0.0)
    if name == "linear":  # :21-25
        return fac * (1 - qs)
    if name == "peak":  # :26-30
        return fac * (1 - 2 * qs + q)
    if name == "cubic_grad":  # :31-39
        return fac * 4 / 3 * np.where(q <= 1, np.where(qs <= 0.5, 18 * q - 12 * qs, -6 * (1 - qs) ** 2), 0.0)
    raise NotImplementedError(name)


# ----------------------------------------------------------------------------------------------------------
# fixed radius search -- open3d.ml FixedRadiusSearch, call site utils/convolutions.py:354-358
# ----------------------------------------------------------------------------------------------------------
def dist2_f32(p, q):
    """Squared L2 distance with the float32 operation order the whole repo agrees on:
    d2 = (dx*dx + dy*dy) + dz*dz, every operation rounded to float32, no FMA contraction."""
    p = np.asarray(p, dtype=F32)
    q = np.asarray(q, dtype=F32)
    d = (p - q).astype(F32)
    xx = (d[..., 0] * d[..., 0]).astype(F32)
    yy = (d[..., 1] * d[..., 1]).astype(F32)
    zz = (d[..., 2] * d[..., 2]).astype(F32)
    return ((xx + yy).astype(F32) + zz).astype(F32)


def fixed_radius_search(points, queries, radius, ignore_query_point=False, return_distances=True):
    """CSR neighbour lists of ``queries`` in ``points`` within ``radius`` (L2, inclusive ``<=``).

    Returns (neighbors_index int32 [P], neighbors_row_splits int64 [Nq+1], neighbors_distance float32 [P])
    with *squared* distances, rows ordered by ascending point index.  ``ignore_query_point`` drops points whose
    position equals the query position in all three coordinates (position equality, not index identity).
    Threshold: ``d2 <= float32(radius)*float32(radius)`` in float32.
    """
    points = np.ascontiguousarray(points, dtype=F32).reshape(-1, 3)
    queries = np.ascontiguousarray(queries, dtype=F32).reshape(-1, 3)
    r = F32(radius)
    thr = F32(r * r)
    nq = queries.shape[0]
    if points.shape[0] == 0 or nq == 0:
        return np.zeros(0, np.int32), np.zeros(nq + 1, np.int64), np.zeros(0, F32)
    tree = cKDTree(points.astype(np.float64))
    # slightly enlarged candidate radius, exact float32 test afterwards
    cand = tree.query_ball_point(queries.astype(np.float64), float(r) * (1 + 1e-4) + 1e-12, return_sorted=True)
    counts = np.fromiter((len(c) for c in cand), dtype=np.int64, count=nq)
    flat = np.fromiter((i for c in cand for i in c), dtype=np.int64, count=int(counts.sum()))
    qidx = np.repeat(np.arange(nq), counts)
    d2 = dist2_f32(points[flat], queries[qidx])
    keep = d2 <= thr
    if ignore_query_point:
        keep &= ~np.all(points[flat] == queries[qidx], axis=1)
    flat, qidx, d2 = flat[keep], qidx[keep], d2[keep]
    row_counts = np.bincount(qidx, minlength=nq).astype(np.int64)
    splits = np.zeros(nq + 1, np.int64)
    np.cumsum(row_counts, out=splits[1:])
    return flat.astype(np.int32), splits, (d2.astype(F32) if return_distances else np.zeros(0, F32))


def fixed_radius_search_bruteforce(points, queries, radius, ignore_query_point=False):
    """O(N*M) definition, used to pin ``fixed_radius_search`` on small inputs."""
    points = np.asarray(points, F32).reshape(-1, 3)
    queries = np.asarray(queries, F32).reshape(-1, 3)
    thr = F32(F32(radius) * F32(radius))
    idx, splits, dist = [], [0], []
    for q in queries:
        d2 = dist2_f32(points, q[None, :])
        m = d2 <= thr
        if ignore_query_point:
            m &= ~np.all(points == q[None, :], axis=1)
        sel = np.nonzero(m)[0]
        idx.append(sel)
        dist.append(d2[sel])
        splits.append(splits[-1] + len(sel))
    return (np.concatenate(idx).astype(np.int32) if idx else np.zeros(0, np.int32),
            np.asarray(splits, np.int64),
            np.concatenate(dist).astype(F32) if dist else np.zeros(0, F32))


# ----------------------------------------------------------------------------------------------------------
# coordinate mapping / interpolation -- open3d.ml continuous_conv
# ----------------------------------------------------------------------------------------------------------
def _sphere_to_cylinder(x, y, z):
    sq = x * x + y * y + z * z
    n = np.sqrt(sq)
    xy2 = x * x + y * y
    with np.errstate(divide="ignore", invalid="ignore"):
        s_cap = np.sqrt(3 * n / (n + np.abs(z)))
        s_side = n / np.sqrt(xy2)
    zero = sq < 1e-12
    cap = (~zero) & (1.25 * z * z > xy2)
    s = np.where(cap, np.nan_to_num(s_cap), np.nan_to_num(s_side, posinf=0.0))
    ox = np.where(zero, 0.0, x * s)
    oy = np.where(zero, 0.0, y * s)
    oz = np.where(zero, 0.0, np.where(cap, np.copysign(n, z), z * 1.5))
    return ox, oy, oz


def _cylinder_to_cube(x, y, z):
    sq = x * x + y * y
    n = np.sqrt(sq)
    zero = sq < 1e-12
    xbig = (~zero) & (np.abs(y) <= np.abs(x))
    with np.errstate(divide="ignore", invalid="ignore"):
        tx = np.copysign(n, x)
        ty = np.copysign(n, y)
        ay = tx * (4 / np.pi) * np.arctan(y / x)
        ax = ty * (4 / np.pi) * np.arctan(x / y)
    ox = np.where(zero, 0.0, np.where(xbig, tx, ax))
    oy = np.where(zero, 0.0, np.where(xbig, ay, ty))
    ox = np.where(np.isfinite(ox), ox, 0.0)
    oy = np.where(np.isfinite(oy), oy, 0.0)
    return ox, oy, z


def map_coordinates(rel, extent, mapping):
    """rel = x_neighbour - x_out [P,3]; returns cube coords in [-0.5,0.5]^3."""
    rel = np.asarray(rel, dtype=np.float64)
    inv = 1.0 / np.float64(extent)
    if mapping == "identity":
        return rel * inv
    u = rel * (2.0 * inv)
    x, y, z = u[:, 0].copy(), u[:, 1].copy(), u[:, 2].copy()
    if mapping == "ball_to_cube_radial":
        rad = np.sqrt(x * x + y * y + z * z)
        amax = np.maximum(np.abs(x), np.maximum(np.abs(y), np.abs(z)))
        with np.errstate(divide="ignore", invalid="ignore"):
            s = np.where(amax < 1e-8, 0.0, 0.5 * rad / amax)
        s = np.where(np.isfinite(s), s, 0.0)
        return np.stack([x * s, y * s, z * s], axis=1)
    if mapping == "ball_to_cube_volume_preserving":
        x, y, z = _sphere_to_cylinder(x, y, z)
        x, y, z = _cylinder_to_cube(x, y, z)
        return 0.5 * np.stack([x, y, z], axis=1)
    raise ValueError(mapping)


def filter_coordinates(c, kernel_size, align_corners, offset=(0.0, 0.0, 0.0)):
    """Cube coords -> continuous filter-array coords g=(gx,gy,gz). kernel_size is [kz,ky,kx] (SURVEY A)."""
    fs = np.array([kernel_size[2], kernel_size[1], kernel_size[0]], dtype=np.float64)  # (x,y,z)
    off = np.asarray(offset, dtype=np.float64)
    if align_corners:
        return (c + 0.5) * (fs - 1) + off
    return (c + 0.5) * fs - 0.5 + off


def interpolation_weights(g, kernel_size, interpolation):
    """Returns (cell [P,C] int64 linear index z*ky*kx + y*kx + x, weight [P,C]) with C = 8 (linear,
    linear_border) or 1 (nearest_neighbor).  SURVEY A.3."""
    kz, ky, kx = (int(v) for v in kernel_size)
    fs = np.array([kx, ky, kz])
    if interpolation == "nearest_neighbor":
        i = np.clip(np.round(g).astype(np.int64), 0, fs - 1)
        cell = (i[:, 2] * ky + i[:, 1]) * kx + i[:, 0]
        return cell[:, None], np.ones((g.shape[0], 1))
    if interpolation == "linear":
        i0 = np.clip(np.trunc(g).astype(np.int64), 0, fs - 1)
        i1 = np.clip(i0 + 1, 0, fs - 1)
        a = np.clip(g - i0, 0.0, 1.0)
        valid0 = valid1 = np.ones_like(a, dtype=bool)
    elif interpolation == "linear_border":
        i0f = np.floor(g)
        a = g - i0f
        i0 = i0f.astype(np.int64)
        i1 = i0 + 1
        valid0 = (i0 >= 0) & (i0 <= fs - 1)
        valid1 = (i1 >= 0) & (i1 <= fs - 1)
        i0 = np.clip(i0, 0, fs - 1)
        i1 = np.clip(i1, 0, fs - 1)
    else:
        raise ValueError(interpolation)
    cells, ws = [], []
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                ix = i1[:, 0] if dx else i0[:, 0]
                iy = i1[:, 1] if dy else i0[:, 1]
                iz = i1[:, 2] if dz else i0[:, 2]
                wx = a[:, 0] if dx else 1 - a[:, 0]
                wy = a[:, 1] if dy else 1 - a[:, 1]
                wz = a[:, 2] if dz else 1 - a[:, 2]
                v = ((valid1[:, 0] if dx else valid0[:, 0]) & (valid1[:, 1] if dy else valid0[:, 1]) &
                     (valid1[:, 2] if dz else valid0[:, 2]))
                cells.append((iz * ky + iy) * kx + ix)
                ws.append(wx * wy * wz * v)
    return np.stack(cells, axis=1), np.stack(ws, axis=1)


# ----------------------------------------------------------------------------------------------------------
# continuous_conv -- open3d.ml.tf.ops.continuous_conv, call sites utils/convolutions.py:431,454,1054
# ----------------------------------------------------------------------------------------------------------
def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, align_corners=True,
                    coordinate_mapping="ball_to_cube_radial", normalize=False, interpolation="linear"):
    """out[o] = sum_n a_n s_n (f[idx_n] @ W(x[idx_n]-y[o]))  (/ sum a_n if normalize)   (SURVEY A.4).

    filters [kz,ky,kx,Cin,Cout]; kwargs exactly as assembled at utils/convolutions.py:414-429."""
    filters = np.asarray(filters, dtype=np.float64)
    kz, ky, kx, cin, cout = filters.shape
    K = kz * ky * kx
    out_positions = np.asarray(out_positions, np.float64).reshape(-1, 3)
    inp_positions = np.asarray(inp_positions, np.float64).reshape(-1, 3)
    inp_features = np.asarray(inp_features, np.float64).reshape(inp_positions.shape[0], cin)
    idx = np.asarray(neighbors_index, np.int64)
    splits = np.asarray(neighbors_row_splits, np.int64)
    n_out = out_positions.shape[0]
    P = idx.shape[0]
    extent = float(np.asarray(extents, np.float64).reshape(-1)[0])
    counts = np.diff(splits)
    o = np.repeat(np.arange(n_out), counts)
    a = np.ones(P) if neighbors_importance is None or len(neighbors_importance) == 0 \
        else np.asarray(neighbors_importance, np.float64)
    s = np.ones(P) if inp_importance is None or len(inp_importance) == 0 \
        else np.asarray(inp_importance, np.float64)[idx]
    rel = inp_positions[idx] - out_positions[o]
    c = map_coordinates(rel, extent, coordinate_mapping)
    g = filter_coordinates(c, (kz, ky, kx), align_corners, offset if offset is not None else (0, 0, 0))
    cell, w = interpolation_weights(g, (kz, ky, kx), interpolation)
    C = cell.shape[1]
    # patch[o, cell, ci] = sum_n a s w f   via a sparse (n_out*K) x N_in operator
    rows = (o[:, None] * K + cell).reshape(-1)
    cols = np.repeat(idx, C)
    vals = (w * (a * s)[:, None]).reshape(-1)
    S = sparse.csr_matrix((vals, (rows, cols)), shape=(n_out * K, inp_positions.shape[0]))
    patch = (S @ inp_features).reshape(n_out, K * cin)
    out = patch @ filters.reshape(K * cin, cout)
    if normalize:
        if neighbors_importance is None or len(neighbors_importance) == 0:
            norm = counts.astype(np.float64)
        else:
            norm = np.bincount(o, weights=a, minlength=n_out)
        nz = norm != 0
        out[nz] /= norm[nz, None]
    return out


def symmetric_kernel(kernel, sym_axis):
    """Effective antisymmetric kernel: concat([-K[::-1,::-1,::-1], K], axis=sym_axis)
    (utils/convolutions.py:410-412)."""
    kernel = np.asarray(kernel)
    return np.concatenate([-kernel[::-1, ::-1, ::-1], kernel], axis=sym_axis)


def cconv_layer(inp_features, inp_positions, out_positions, extents, kernel, bias=None, *,
                align_corners=True, coordinate_mapping="ball_to_cube_radial", interpolation="linear",
                normalize=True, ignore_query_points=False, window_name=None, window_fac=1.0,
                symmetric=False, sym_axis=2, return_nns=False, backend=None):
    """ContinuousConv.call (utils/convolutions.py:277-470) for scalar extents, no dense-for-center,
    non-circular kernels, linear activation."""
    B = backend if backend is not None else _Backend64
    dt = B.dtype
    radius = F32(0.5) * F32(extents)  # :353
    idx, splits, d2 = B.fixed_radius_search(inp_positions, out_positions, radius,
                                            ignore_query_point=ignore_query_points,
                                            return_distances=window_name is not None)
    if window_name is not None:
        q = d2.astype(dt) / (dt(radius) * dt(radius))  # :361-362
        imp = B.window(window_name, q, window_fac)  # :378-379
    else:
        imp = None
    k = np.asarray(kernel, dt)
    if symmetric:
        k = symmetric_kernel(k, sym_axis)
    kw = dict(out_positions=out_positions, extents=extents, offset=(0, 0, 0), inp_positions=inp_positions,
              inp_importance=None, neighbors_index=idx, neighbors_importance=imp,
              neighbors_row_splits=splits, align_corners=align_corners,
              coordinate_mapping=coordinate_mapping, normalize=normalize, interpolation=interpolation)
    feats = np.asarray(inp_features, dt)
    out = B.continuous_conv(filters=k, inp_features=feats, **kw)  # :431
    if symmetric:  # :433-458
        assert inp_positions.shape == out_positions.shape
        wk = k.reshape(k.shape[0], k.shape[1], k.shape[2], 1, -1)
        w_values = B.continuous_conv(filters=wk, inp_features=np.ones((feats.shape[0], 1), dt), **kw)
        res = w_values.reshape(-1, k.shape[-2], k.shape[-1])
        out = out + np.einsum("nc,nco->no", feats, res)
    if bias is not None:
        out = out + np.asarray(bias, dt)
    if return_nns:
        return out, (idx, splits, d2)
    return out


class _Backend64:
    """float64 NumPy ops (the default backend); oracle/o32.py provides the float32 C one."""
    dtype = np.float64
    fixed_radius_search = staticmethod(fixed_radius_search)
    continuous_conv = staticmethod(continuous_conv)
    window = staticmethod(window)

    @staticmethod
    def dense(x, kernel, bias):
        return dense(x, kernel, bias)


def point_sampling(inp_features, inp_positions, out_positions, extents, window_name=None, normalize=True):
    """PointSampling.call (utils/convolutions.py:928-1058): 1x1x1 identity kernel, default mapping/interp."""
    cin = np.asarray(inp_features).shape[-1]
    kernel = np.eye(cin).reshape(1, 1, 1, cin, cin)
    return cconv_layer(inp_features, inp_positions, out_positions, extents, kernel, None,
                       align_corners=True, coordinate_mapping="ball_to_cube_radial", interpolation="linear",
                       normalize=normalize, ignore_query_points=False, window_name=window_name)


# ----------------------------------------------------------------------------------------------------------
# multi-scale sampling -- utils/tools/losses.py:136-181, 249-284
# ----------------------------------------------------------------------------------------------------------
def compute_density(out_pos, in_pos, radius, window_name=None):
    """utils/tools/losses.py:287-308: sum over neighbours (self included) of win(d^2/r^2); identity window if None."""
    idx, splits, d2 = fixed_radius_search(np.asarray(in_pos, F32), np.asarray(out_pos, F32), F32(radius))
    q = d2.astype(np.float64) / (np.float64(F32(radius)) ** 2)
    w = window(window_name, q) if window_name is not None else q
    rows = np.repeat(np.arange(len(splits) - 1), np.diff(splits))
    return np.bincount(rows, weights=w, minlength=len(splits) - 1)


def grid_pos(pos, voxel_size, centralize=False, pad=0, hyst=0.1):
    """Lattice points of the voxel grid touched by any particle (utils/tools/losses.py:136-181).
    float32 arithmetic like the reference; first-occurrence order like ``tf.unique``."""
    pos = np.asarray(pos, F32)
    v = np.asarray(voxel_size, F32).reshape(3)
    center = None
    if centralize:  # :137-139
        center = pos.astype(np.float64).mean(axis=0).astype(F32)  # float64 accumulate, rounded once (repo convention)
        pos = (pos - center).astype(F32)
    vm = np.maximum(v, F32(1e-5))
    h = np.where(v >= 1e-5, F32(hyst), F32(0.0)).astype(F32)
    scaled = (pos / vm).astype(F32)
    dpos = np.concatenate([np.floor((scaled - h).astype(F32)).astype(np.int32),
                           np.floor((scaled + h).astype(F32)).astype(np.int32)], axis=0)  # :142-150
    rng = [np.arange(-pad, 2 + pad) if v[i] >= 1e-5 else np.arange(0, 1) for i in range(3)]  # :151-160
    off = np.stack(np.meshgrid(*rng, indexing="ij"), axis=-1).reshape(1, -1, 3)
    dpos = (dpos[:, None, :] + off).reshape(-1, 3)  # :163-164
    minp = dpos.min(axis=0)  # :167
    maxp = dpos.max(axis=0) - minp + 1
    lin = ((dpos - minp).astype(np.int64) * np.array([1, maxp[0], maxp[0] * maxp[1]], np.int64)).sum(axis=-1)
    _, first = np.unique(lin, return_index=True)
    uniq = lin[np.sort(first)]  # tf.unique keeps first-occurrence order (:171)
    gpos = np.stack([uniq % maxp[0], uniq // maxp[0] % maxp[1], uniq // (maxp[0] * maxp[1])], axis=-1) + minp
    if centralize:  # :176-179
        return (gpos.astype(F32) * v + center).astype(F32)
    return (gpos.astype(F32) * v + v / F32(2)).astype(F32)


def get_dilated_pos(pos, strides, voxel_size=None, centralize=False, pad=0, hyst=0.1, return_idx=False):
    """utils/tools/losses.py:249-284: voxel lattices, or (voxel_size None) nested farthest-point subsets of
    N // stride points taken from the previous scale (:274-282); idx[s] = indices of scale s inside scale s-1."""
    out, idx = [], []
    for s in strides:
        if s == 1:
            out.append(np.asarray(pos, F32))
            idx.append(None)
        elif voxel_size is None:
            from . import pointset
            cnt = max(len(pos) // int(s), 1)
            idx.append(pointset.farthest_point_sample(cnt, out[-1]))
            out.append(out[-1][idx[-1]])
        else:
            out.append(grid_pos(pos, np.asarray(voxel_size, F32) * F32(s), centralize, pad, hyst))
    return (out, idx) if return_idx else out


# ----------------------------------------------------------------------------------------------------------
# model -- models/pbf_model.py, models/hrnet.py, models/sym_net.py, models/cconv.py
# ----------------------------------------------------------------------------------------------------------
def dense(x, kernel, bias):
    """Keras Dense, linear: x @ W + b (W is [Cin,Cout])."""
    return np.asarray(x, np.float64) @ np.asarray(kernel, np.float64) + np.asarray(bias, np.float64)


def align_vector(v0, v1):
    """models/pbf_model.py:12-28."""
    v0 = np.asarray(v0, np.float64)
    v1 = np.asarray(v1, np.float64)
    a = v0 / (np.linalg.norm(v0) + 1e-9)
    b = v1 / (np.linalg.norm(v1) + 1e-9)
    v = np.cross(a, b)
    c = float(a @ b)
    s = np.linalg.norm(v)
    if s < 1e-6:
        return np.eye(3) * (-1.0 if c < 0 else 1.0)
    vx = np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])
    return np.eye(3) + vx + vx @ vx / (1 + c)


class ModelO64:
    """Reference dataflow of PBFNet/HRNet/SymNet/CConv on a ``weights`` dict keyed by checkpoint-style names
    (SURVEY Appendix B).  ``cfg`` is the ``model:`` section of a reference YAML."""

    def __init__(self, cfg, weights, backend=None):
        self.B = backend if backend is not None else _Backend64
        d = dict(kernel_size=[4, 4, 4], strides=[1], particle_radii=[0.05],
                 coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", window=None,
                 ignore_query_points=False, grav=-9.81, transformation={}, timestep=0.01, use_vel=True,
                 use_acc=True, use_box_feats=True, use_bnds=True, voxel_size=None, centralize=False,
                 out_scale=[0.01, 0.01, 0.01], sample_pad=0, sample_hyst=0.1, part_scale=1.0,
                 add_merge=False, sym_kernel_size=[6, 6, 6], sym_axis=2, window_sym=None, window_dens=None,
                 dens_feats=False, pres_feats=False, dens_norm=False, use_pre_adv=False, use_feats=False,
                 dens_radius=None, rest_dens=3.5, stiffness=20.0)
        d.update(cfg)
        self.c = d
        self.w = weights
        self.name = d.get("name", "SymNet")
        self.nns_fluid = None

    # -- helpers ------------------------------------------------------------------------------------------
    def _conv(self, key, feats, inp_pos, out_pos, extent, ignore_q=None, window=None, **kw):
        c = self.c
        return cconv_layer(feats, inp_pos, out_pos, extent, self.w[key + "/kernel"],
                           self.w.get(key + "/bias"), align_corners=True,
                           coordinate_mapping=c["coordinate_mapping"], interpolation=c["interpolation"],
                           normalize=False,
                           ignore_query_points=c["ignore_query_points"] if ignore_q is None else ignore_q,
                           window_name=c["window"] if window is None else window, backend=self.B, **kw)

    def _dense(self, key, x):
        return self.B.dense(x, self.w[key + "/kernel"], self.w[key + "/bias"])

    # -- BaseModel.call: models/base_model.py:23-29 -----------------------------------------------------------
    def __call__(self, pos, vel, acc, box, bfeats, feats=None):
        c = self.c
        pos = np.asarray(pos, np.float64); vel = np.asarray(vel, np.float64)
        acc = None if acc is None else np.asarray(acc, np.float64)
        box = np.asarray(box, np.float64); bfeats = np.asarray(bfeats, np.float64)
        pos0, vel0 = pos, vel
        tr = c["transformation"] or {}
        # transform: models/pbf_model.py:252-280
        if "translate" in tr:
            pos = pos + np.asarray(tr["translate"]); box = box + np.asarray(tr["translate"])
        if "scale" in tr:
            sc = np.asarray(tr["scale"], np.float64)
            pos = pos * sc; box = box * sc; vel = vel * sc
            if acc is not None:
                acc = acc * sc
        R = None
        if "grav_eqvar" in tr:
            R = align_vector(np.asarray(tr["grav_eqvar"], np.float64), acc[0])
            pos, vel, acc, box, bfeats = pos @ R, vel @ R, acc @ R, box @ R, bfeats @ R
        tpos, tvel, tacc = pos, vel, acc
        # preprocess: models/pbf_model.py:303-438
        dt = c["timestep"]
        a_int = acc if acc is not None else np.array([0.0, c["grav"], 0.0])
        # the integration itself runs in float32 with separately rounded ops, like the reference's TF ops (:234-240)
        vel2 = (vel.astype(F32) + (F32(dt) * np.asarray(a_int, F32)).astype(F32)).astype(F32)
        pos2 = (pos.astype(F32) + (F32(dt) * vel2).astype(F32)).astype(F32)
        vel2, pos2 = vel2.astype(np.float64), pos2.astype(np.float64)
        ext = np.asarray(c["particle_radii"], F32) * F32(2)
        lo = pos2.min(axis=0) - float(ext[-1]); hi = pos2.max(axis=0) + float(ext[-1])
        f = np.all((box >= lo) & (box <= hi), axis=1)  # :330-334
        box_c, bfeats_c = box[f], bfeats[f]
        ff = [np.ones((pos2.shape[0], 1))]
        if c["use_vel"]:
            ff.append(vel2)
        if c["use_acc"]:
            ff.append(acc)
        if c["use_feats"]:
            ff.append(np.asarray(feats, np.float64))
        bf = [np.ones((box_c.shape[0], 1))]
        if c["use_box_feats"]:
            bf.append(bfeats_c)
        all_pos = np.concatenate([pos2, box_c], axis=0).astype(F32)
        pos2_32 = pos2.astype(F32); box_32 = box_c.astype(F32)
        n_f = pos2.shape[0]
        dens_radius = c["dens_radius"] if c["dens_radius"] is not None else c["particle_radii"]
        dens0 = None
        if c["dens_feats"] or c["dens_norm"] or c["pres_feats"]:  # models/pbf_model.py:351-367
            dens0 = compute_density(all_pos, all_pos, dens_radius[0], c["window_dens"])
            if c["dens_feats"]:
                ff.append(dens0[:n_f, None]); bf.append(dens0[n_f:, None])
            if c["pres_feats"]:  # utils/tools/losses.py:367-377
                pres = np.maximum(c["stiffness"] * ((dens0 / c["rest_dens"]) ** 7 - 1), 0.0)
                ff.append(pres[:n_f, None]); bf.append(pres[n_f:, None])
        fluid_feats = np.concatenate(ff, axis=1)
        box_feats = np.concatenate(bf, axis=1)
        ps = c["part_scale"]
        ans_conv, self.nns_fluid = self._conv("fluid_convs", fluid_feats * ps, pos2_32, all_pos, ext[0],
                                              return_nns=True)  # :378
        ans_dense = self._dense("fluid_dense", fluid_feats)
        ans_obs = self._conv("obs_convs", box_feats * ps, box_32, all_pos, ext[0])  # :382
        ans_dense_obs = self._dense("obs_dense", box_feats)
        ans_dense = np.concatenate([ans_dense, ans_dense_obs], axis=0)
        if c["use_pre_adv"]:  # :388-399
            pre = np.ones((n_f, 1))
            if c["use_vel"]:
                pre = np.concatenate([pre, vel], axis=1)
            ans_adv = self._conv("adv_convs/0", pre * ps, pos.astype(F32), all_pos, ext[0])
            ans_dens_adv = np.concatenate([self._dense("adv_dense/0", pre), ans_dense_obs], axis=0)
            feats = np.concatenate([ans_conv, ans_obs, ans_adv, ans_dense, ans_dens_adv], axis=1)
        else:
            feats = np.concatenate([ans_conv, ans_obs, ans_dense], axis=1)
        dil, self.fps_idx = get_dilated_pos(all_pos if c["use_bnds"] else pos2_32, c["strides"], c["voxel_size"],
                                            c["centralize"], c["sample_pad"], c["sample_hyst"], return_idx=True)  # :413-419
        self.dilated_pos = dil
        self.dens = None
        if c["dens_norm"]:  # :421-435 (the radius is passed as the sampling extent)
            self.dens = [(dens0 if c["use_bnds"] else dens0[:n_f])[:, None]]
            for sc in range(1, len(dens_radius)):
                d = point_sampling(self.dens[-1], dil[sc - 1], dil[sc], F32(dens_radius[sc]), c["window_dens"], True)
                self.dens.append(np.maximum(d, 1e-2))
        # forward
        if self.name == "CConv":
            out = self._forward_cconv(dil, feats, ext, n_f)
        else:
            out = self._forward_hrnet(dil, feats, ext, n_f)
            if self.name == "SymNet":
                out = self._forward_sym(out, feats, all_pos, ext, n_f)
        self.net_out = out
        # postprocess: models/pbf_model.py:440-489
        if out.shape[1] == 1:
            out = np.repeat(out, 3, axis=1)
        elif out.shape[1] == 2:
            out = np.concatenate([out, out[:, :1]], axis=1)
        corr = np.asarray(c["out_scale"], np.float64) * out[:n_f]
        self.pos_correction = corr
        pos_new = pos2 + corr
        vel_new = (pos_new - tpos) / dt  # :242-250
        # inv_transform: :282-301
        if R is not None:
            pos_new, vel_new = pos_new @ R.T, vel_new @ R.T
        if "scale" in tr:
            sc = np.maximum(np.asarray(tr["scale"], np.float64), 1e-5)
            pos_new, vel_new = pos_new / sc, vel_new / sc
        if "translate" in tr:
            pos_new = pos_new - np.asarray(tr["translate"])
        return pos_new, vel_new

    # -- models/hrnet.py:69-133 (k == 0 convs only) --------------------------------
    def _forward_hrnet(self, pos, feats, ext, n_f):
        c = self.c
        lc = c["layer_channels"]
        if self.name == "SymNet":
            lc = lc[:-1]
        if not c["use_bnds"]:
            feats = feats[:n_f]
        n = 4 if c["use_pre_adv"] else 2  # _all_convs numbering: 0 fluid_obs, 1 obs_conv (, 2-3 adv_conv0/1) (models/pbf_model.py:223)
        ans_convs = [[feats]]
        for i in range(1, len(lc)):
            ans = []
            for j in range(len(lc[i])):
                imp = c["part_scale"] if j == 0 else 1.0
                inp = []
                if len(lc[i][j]) != 1:
                    raise NotImplementedError("k>0 convs per scale")
                for l in range(len(lc[i - 1])):
                    fe = np.maximum(ans_convs[-1][l], 0.0)
                    e = ext[max(l, j)]
                    if c["dens_norm"] and l < len(self.dens):  # models/hrnet.py:87-89
                        fe = np.concatenate([fe, fe / self.dens[l] ** 2], axis=1)
                    key = "_all_convs/%d" % n
                    n += 1
                    a = self._conv(key, fe * imp, pos[l], pos[j], e,
                                   ignore_q=c["ignore_query_points"] and (j == l))
                    if j == l:
                        a = a + self._dense("denses/%d/%d/0/%d" % (i - 1, j, l), fe)
                        if a.shape[1] == ans_convs[-1][j].shape[1]:
                            a = a + ans_convs[-1][j]
                    elif c["voxel_size"] is None:  # models/hrnet.py:100-113: nested farthest-point subsets
                        key_d = "denses/%d/%d/0/%d" % (i - 1, j, l)
                        if j > l:
                            for t in range(l, j):
                                fe = fe[self.fps_idx[t + 1]]
                            a = a + self._dense(key_d, fe)
                        else:
                            ind = self.fps_idx[j + 1]
                            for t in range(j + 1, l):
                                ind = ind[self.fps_idx[t + 1]]
                            a = a.copy()
                            np.add.at(a, ind, self._dense(key_d, fe))
                    inp.append(a)
                ans.append(sum(inp) if c["add_merge"] else np.concatenate(inp, axis=1))
            ans_convs.append(ans)
        return ans_convs[-1][0]

    # -- models/sym_net.py:55-69 -----------------------------------------------------------------------------
    def _forward_sym(self, ans, feats, all_pos, ext, n_f):
        c = self.c
        if not c["use_bnds"]:
            ans = np.concatenate([ans, feats[n_f:]], axis=0)
        n_sym = len(c["layer_channels"][-1][-1])
        for i in range(n_sym):
            ans = np.maximum(ans, 0.0)
            ans = cconv_layer(ans * c["part_scale"], all_pos, all_pos, ext[0], self.w["sym_convs/%d/kernel" % i],
                              None, align_corners=True, coordinate_mapping=c["coordinate_mapping"],
                              interpolation=c["interpolation"], normalize=False, ignore_query_points=True,
                              window_name=c["window_sym"], symmetric=True, sym_axis=c["sym_axis"], backend=self.B)
        return ans

    # -- models/cconv.py:50-69 -------------------------------------------------------------------------------
    def _forward_cconv(self, pos, feats, ext, n_f):
        c = self.c
        p = pos[0]  # models/cconv.py:52-53
        feats = feats[:p.shape[0]]
        ans = feats
        for i in range(1, len(c["layer_channels"])):
            fe = np.maximum(ans, 0.0)
            a = self._conv("_all_convs/%d" % (i + 1), fe, p, p, ext[0]) + self._dense("denses/%d" % (i - 1), fe)
            if a.shape[1] == ans.shape[1]:
                a = a + ans
            ans = a
        return ans
