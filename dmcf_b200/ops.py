"""Torch-tensor front end of the C ABI (include/dmcf_b200.h): the ops DMCF reaches through ``open3d.ml.tf``.

Mirrors the reference's op surface for the hot path:
  * ``fixed_radius_search``  <- ml3d.layers.FixedRadiusSearch (utils/convolutions.py:207-210, 354-358)
  * ``continuous_conv``      <- ml3d.ops.continuous_conv       (utils/convolutions.py:414-431)
  * ``reduce_subarrays_sum`` <- o3dml.ops.reduce_subarrays_sum (models/pbf_model.py:450-453)
  * ``dense``                <- tf.keras.layers.Dense           (models/pbf_model.py:140-152)
Every function takes CUDA float32 tensors, enqueues on the current torch stream and raises on anything else:
there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import threading
from collections import namedtuple

import torch

from . import _lib
from ._lib import ConvDesc, DmcfError, Grid, check

MAPPINGS = {"identity": 0, "ball_to_cube_radial": 1, "ball_to_cube_volume_preserving": 2}
INTERPOLATIONS = {"linear": 0, "linear_border": 1, "nearest_neighbor": 2}
WINDOWS = {None: 0, "poly6": 1, "cubic": 2, "linear": 3, "peak": 4, "cubic_grad": 5}

NeighborSearchResult = namedtuple("NeighborSearchResult",
                                  ["neighbors_index", "neighbors_row_splits", "neighbors_distance"])

MAX_CELLS = 1 << 24

# bench.py sets this to a list to get one {start,end CUDA events + shape metadata} record per conv launch
PROFILE = None

# ---- sync-free steps: capacity-sized buffers, device-side counts, a plan of the data-dependent sizes ---------------------
# A step has a handful of data-dependent sizes (cell-grid extents, neighbour pairs per list, culled boundary rows, lattice
# points per scale).  The reference's graph mode (pipelines/simulator.py:57, tf.function) hides them inside TensorFlow's
# dynamic shapes; here a StepPlan records them once in a MEASURING step (exact sizes read back from the device, i.e. with host
# syncs) and every later step REPLAYS the plan: buffers get the recorded size plus slack, the true counts stay in device
# memory (`with_count`), kernels are launched for the capacity and ignore rows beyond the count, and anything that outgrows
# its capacity raises a flag in `StepPlan.flags` that the caller reads ONCE at the end of the step (dmcf_b200/simulator.py).
# Such a step makes no host sync and can be captured in a CUDA graph.
_TLS = threading.local()  # the active plan is per thread (slab ranks may run as threads of one process)


def _PLAN():
    return getattr(_TLS, "plan", None)


def get_plan():
    """The StepPlan the calling thread is measuring / replaying under, or None (eager)."""
    return getattr(_TLS, "plan", None)


def set_plan(plan):
    _TLS.plan = plan


class StepPlan:
    """Sequence of the data-dependent size decisions of one model step, in call order (the control flow of a step is fixed by
    the model, so the i-th decision of every step is the same one)."""

    PAIR_SLACK, ROW_SLACK, GRID_PAD_CELLS, GRID_PAD_FRAC = 1.25, 1.25, 3, 0.08

    def __init__(self, device):
        self.entries = []
        self.mode = "measure"
        self.i = 0
        self.device = device
        self.flags = None  # int32 [2 * n_entries]: hard (results invalid) flags first, soft (re-plan, results fine) after

    def begin(self, mode):
        self.mode, self.i = mode, 0
        if mode == "measure":
            self.entries = []
            self.flags = None
        else:
            if self.flags is None:
                self.flags = torch.zeros(2 * max(len(self.entries), 1), dtype=torch.int32, device=self.device)
            else:
                self.flags.zero_()

    def record(self, kind, **vals):
        self.entries.append(dict(kind=kind, **vals))

    def next(self, kind):
        if self.i >= len(self.entries) or self.entries[self.i]["kind"] != kind:
            raise DmcfError(f"step plan out of sequence at entry {self.i}: wanted {kind!r}, plan has "
                            f"{self.entries[self.i]['kind'] if self.i < len(self.entries) else 'nothing'!r}")
        e = self.entries[self.i]
        slot = self.i
        self.i += 1
        return e, slot

    def hard(self, slot):
        """int32 view of the hard-overflow flag of entry ``slot``."""
        return self.flags[slot:slot + 1]

    def soft(self, slot):
        return self.flags[len(self.entries) + slot:len(self.entries) + slot + 1]


class no_plan:
    """Context: run ops outside the plan (static, cached pieces such as the cell order of the boundary)."""

    def __enter__(self):
        self.saved = get_plan()
        set_plan(None)

    def __exit__(self, *exc):
        set_plan(self.saved)


def with_count(t, n_dev):
    """Marks ``t`` ([capacity, ...]) as holding only ``n_dev`` (int32 device tensor [1]) valid leading rows."""
    t._dmcf_n_dev = n_dev
    return t


def count_of(t):
    return getattr(t, "_dmcf_n_dev", None)


def valid_rows_mask(t):
    """Bool [capacity] mask of the valid rows of a capacity-sized tensor (all True without a device-side count)."""
    n_dev = count_of(t)
    if n_dev is None:
        return None
    return torch.arange(t.shape[0], device=t.device, dtype=torch.int32) < n_dev


def compact_mask(mask, capacity, overflow=None):
    """Stable compaction of the True rows of ``mask`` without a host sync: (index int64 [capacity] -- entries beyond the count
    are 0 --, count int32 [1]).  More than ``capacity`` True rows set ``overflow``."""
    n = mask.shape[0]
    c = torch.cumsum(mask.to(torch.int32), 0, dtype=torch.int32)
    count = c[-1:].clone() if n > 0 else torch.zeros(1, dtype=torch.int32, device=mask.device)
    dst = torch.where(mask & (c <= capacity), (c - 1).to(torch.int64), torch.full((), capacity, dtype=torch.int64, device=mask.device))
    idx = torch.zeros(capacity + 1, dtype=torch.int64, device=mask.device)
    idx.scatter_(0, dst, torch.arange(n, device=mask.device, dtype=torch.int64))
    if overflow is not None:
        overflow.copy_(torch.maximum(overflow, (count > capacity).to(torch.int32)))
    return idx[:capacity], torch.clamp(count, max=capacity)


def _prof_begin(kind, **meta):
    """bench.py's per-op timing hook for the HBM-bound ops around the convs (record has a ``kind`` key; conv records do not)."""
    if PROFILE is None:
        return None
    rec = dict(kind=kind, start=torch.cuda.Event(enable_timing=True), end=torch.cuda.Event(enable_timing=True), **meta)
    rec["start"].record()
    return rec


def _prof_end(rec):
    if rec is not None:
        rec["end"].record()
        PROFILE.append(rec)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _req(t, name, dtype=torch.float32, dim=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise DmcfError(f"{name} must be a CUDA tensor: dmcf_b200 has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if dim is not None and t.dim() != dim:
        raise ValueError(f"{name} must have rank {dim}, got shape {tuple(t.shape)}")
    return t


def _rows(t, name):
    """2-D tensor whose rows are contiguous; returns (tensor, row stride in elements)."""
    _req(t, name, dim=2)
    if t.stride(1) != 1 and t.shape[1] > 1:
        t = t.contiguous()
    if t.shape[0] > 1 and t.stride(0) < t.shape[1]:
        t = t.contiguous()
    return t, (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0)))


def _pos(t, name):
    _req(t, name, dim=2)
    if t.shape[1] != 3:
        raise ValueError(f"{name} must have shape [N,3], got {tuple(t.shape)}")
    return t.contiguous()


def select_rows(mask, *tensors, bounded=True):
    """Rows of every tensor where ``mask`` is True (the mask must already exclude padding rows).  Eager: boolean indexing (a
    host sync).  Under a StepPlan the measured count is recorded and a replay returns capacity-sized tensors (measured count
    plus slack) with the device-side count attached.  ``bounded=False``: do not cap the capacity at the number of mask rows
    (message buffers whose size two ranks must derive from the same measured count)."""
    if _PLAN() is not None and _PLAN().mode == "replay":
        e, slot = _PLAN().next("rows")
        cap = int(e["n"] * StepPlan.ROW_SLACK) + 256
        if bounded:
            cap = min(cap, max(int(mask.shape[0]), 1))
        idx, cnt = compact_mask(mask, cap, _PLAN().hard(slot))
        return [with_count(t[idx], cnt) for t in tensors]
    out = [t[mask] for t in tensors]
    if _PLAN() is not None and _PLAN().mode == "measure":
        _PLAN().record("rows", n=int(out[0].shape[0]))
    return out


def planned_rows(n_measured=None):
    """Capacity of a buffer whose row count another rank decides (halo / migration arrivals): measure mode records the exact
    count and returns it; replay returns (capacity, hard-overflow flag) derived from the recorded count."""
    if _PLAN() is not None and _PLAN().mode == "replay":
        e, slot = _PLAN().next("rows")
        return int(e["n"] * StepPlan.ROW_SLACK) + 256, _PLAN().hard(slot)
    if _PLAN() is not None and _PLAN().mode == "measure":
        _PLAN().record("rows", n=int(n_measured))
    return int(n_measured), None


def append_rows(dst, base, src, index=None, overflow=None, want_count=True):
    """dst[base + i] = src[index[i] if index is given else i] for the valid rows i of ``src`` (``index`` / ``src`` may be capacity
    sized with a count); ``base`` is an int or an int32 device tensor [1].  One kernel (dmcf_rows_append), no host sync.
    Returns the new count (int32 device tensor [1]) when ``want_count``."""
    lib = _lib.load()
    if dst.dtype != torch.float32 or src.dtype != torch.float32 or dst.dim() != 2 or src.dim() != 2:
        raise TypeError("append_rows moves float32 [rows, width] blocks")
    if dst.stride(1) != 1 or (src.shape[1] > 1 and src.stride(1) != 1):
        raise ValueError("append_rows needs unit-stride rows")
    width = int(src.shape[1])
    if dst.shape[1] != width:
        raise ValueError("append_rows: widths differ")
    lead = index if index is not None else src
    n_src = int(lead.shape[0])
    cnt = count_of(lead)
    if index is not None:
        _req(index, "index", torch.int64, 1)
        index = index.contiguous()
    new_cnt = torch.empty(1, dtype=torch.int32, device=dst.device) if want_count else None
    base_dev = base if isinstance(base, torch.Tensor) else None
    check(lib.dmcf_rows_append(_p(dst), dst.stride(0) if dst.shape[0] > 1 else max(width, dst.stride(0)), dst.shape[0],
                               0 if base_dev is not None else int(base), _p(base_dev), _p(src),
                               src.stride(0) if src.shape[0] > 1 else max(width, src.stride(0)), n_src, _p(cnt), _p(index), width,
                               _p(new_cnt), _p(overflow), _stream()))
    return new_cnt


def concat_rows(parts, extra_rows=0):
    """Row-wise concatenation of point sets / feature blocks that may be capacity-sized: the valid rows of every part, in
    order, form the valid prefix of the result (capacity = sum of the capacities, count = sum of the counts, on the device).
    All parts exact -> plain torch.cat."""
    if all(count_of(t) is None for t in parts):
        return torch.cat(parts, dim=0)
    cap = sum(int(t.shape[0]) for t in parts) + int(extra_rows)
    first = parts[0]
    if first.dim() != 2 or first.dtype != torch.float32:
        raise TypeError("concat_rows of capacity-sized parts moves float32 [rows, width] blocks")
    out = torch.zeros((cap, first.shape[1]), dtype=torch.float32, device=first.device)  # padding rows stay finite
    base = 0
    for t in parts:
        if int(t.shape[0]) == 0:
            continue
        if count_of(t) is None and not isinstance(base, torch.Tensor):
            out[base:base + t.shape[0]] = t
            base += int(t.shape[0])
        else:
            base = append_rows(out, base, t if t.stride(-1) == 1 else t.contiguous())
    if not isinstance(base, torch.Tensor):
        base = torch.full((1,), base, dtype=torch.int32, device=first.device)
    return with_count(out, base)


def pad_rows(t, capacity, count=None):
    """``t`` (exact rows, or capacity-sized with a count) in a buffer of ``capacity`` rows with the count on the device."""
    n = int(t.shape[0])
    cnt = count_of(t) if count is None else count
    if cnt is None:
        cnt = torch.full((1,), n, dtype=torch.int32, device=t.device)
    if n == capacity:
        return with_count(t, cnt)
    out = torch.zeros((capacity, *t.shape[1:]), dtype=t.dtype, device=t.device)
    m = min(n, capacity)
    out[:m] = t[:m]
    return with_count(out, torch.clamp(cnt, max=capacity))


def trim(t):
    """Exact-sized copy of a capacity-sized tensor (ONE host sync: reads the count)."""
    cnt = count_of(t)
    if cnt is None:
        return t
    return t[: int(cnt.item())]


class CellList:
    """Uniform cell list over a point set (the reference's ``build_spatial_hash_table`` result)."""

    def __init__(self, points, cell_size, origin=None, dims=None, n_dev=None):
        lib = _lib.load()
        if n_dev is None:
            n_dev = count_of(points)
        points = _pos(points, "points")
        n = points.shape[0]
        if n >= 2 ** 31:
            raise DmcfError("too many points")
        cell_size = float(cell_size)
        if not cell_size > 0:
            raise ValueError("cell_size must be positive")
        box_lo = box_hi = None  # the points' true bounding box when known (occupancy hint for the search kernels)
        if (origin is None or dims is None) and _PLAN() is not None and _PLAN().mode == "replay":
            # planned grid: the measured bounding box padded by a few cells; points that leave it are clamped into the border
            # cells (always correct) and raise the SOFT flag so that the caller re-plans after this step
            e, slot = _PLAN().next("grid")
            box_lo, box_hi = e["lo"], e["hi"]
            pad = [StepPlan.GRID_PAD_CELLS * cell_size + StepPlan.GRID_PAD_FRAC * (e["hi"][a] - e["lo"][a]) for a in range(3)]
            origin = [e["lo"][a] - pad[a] for a in range(3)]
            while True:
                dims = [int((e["hi"][a] + pad[a] - origin[a]) / cell_size) + 1 for a in range(3)]
                if dims[0] * dims[1] * dims[2] <= MAX_CELLS:
                    break
                cell_size *= 1.26
            if n > 0:
                bounds = e.get("_bounds")
                if bounds is None:  # uploaded once per plan (outside any graph capture: the first replay runs eagerly)
                    lo_t = torch.tensor(origin, dtype=torch.float32)
                    hi_t = torch.tensor([origin[a] + dims[a] * cell_size for a in range(3)], dtype=torch.float32)
                    bounds = e["_bounds"] = (lo_t.to(points.device), hi_t.to(points.device))
                out = (points < bounds[0]) | (points > bounds[1])
                m = None if n_dev is None else (torch.arange(n, device=points.device, dtype=torch.int32) < n_dev)
                out = out.any(dim=1) if m is None else (out.any(dim=1) & m)
                soft = _PLAN().soft(slot)
                soft.copy_(torch.maximum(soft, out.any().to(torch.int32).reshape(1)))
        if origin is None or dims is None:
            if n > 0:
                if n_dev is not None:
                    raise DmcfError("a capacity-sized point set needs a planned grid (origin / dims)")
                lo = points.amin(dim=0)
                hi = points.amax(dim=0)
                lohi = torch.stack([lo, hi]).cpu()  # one host sync per build; pass origin/dims (or run under a StepPlan) to avoid it
                lo, hi = lohi[0].tolist(), lohi[1].tolist()
            else:
                lo, hi = [0.0] * 3, [0.0] * 3
            box_lo, box_hi = lo, hi
            if _PLAN() is not None and _PLAN().mode == "measure":
                _PLAN().record("grid", lo=lo, hi=hi)
            while True:
                dims = [int((hi[a] - lo[a]) / cell_size) + 1 for a in range(3)]
                if dims[0] * dims[1] * dims[2] <= MAX_CELLS:
                    break
                cell_size *= 1.26
            origin = lo
        self.points = points
        self.cell_size = cell_size
        self.n_cells = int(dims[0]) * int(dims[1]) * int(dims[2])
        dev = points.device
        self.cell_start = torch.empty(self.n_cells + 1, dtype=torch.int32, device=dev)
        # with a device-side count the entries beyond it are never written: keep them valid indices (callers gather with them)
        self.sorted_index = (torch.zeros if n_dev is not None else torch.empty)(max(n, 1), dtype=torch.int32, device=dev)
        self.sorted_pos = torch.empty((max(n, 1), 4), dtype=torch.float32, device=dev)
        self.n_dev = n_dev
        g = Grid()
        g.origin[:] = [float(v) for v in origin]
        g.inv_cell = 1.0 / cell_size
        g.dims[:] = [int(v) for v in dims]
        g.n_points = n
        g.cell_start = self.cell_start.data_ptr()
        g.sorted_index = self.sorted_index.data_ptr()
        g.sorted_pos = self.sorted_pos.data_ptr()
        g.n_points_dev = None if n_dev is None else n_dev.data_ptr()
        if box_lo is not None and n > 0:
            occupied = 1.0
            for a in range(3):
                occupied *= max(1.0, (box_hi[a] - box_lo[a]) / cell_size)
            g.mean_occupancy = n / occupied
        self.grid = g
        ws_bytes = lib.dmcf_grid_workspace_bytes(n, self.n_cells)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        rec = _prof_begin("cell_list", n_points=n, n_cells=self.n_cells)
        check(lib.dmcf_grid_build(_p(points), C.byref(g), _p(ws), ws_bytes, _stream()))
        _prof_end(rec)


def exclusive_scan(counts, out_dtype=torch.int64):
    lib = _lib.load()
    _req(counts, "counts", torch.int32, 1)
    counts = counts.contiguous()
    n = counts.shape[0]
    out = torch.empty(n + 1, dtype=out_dtype, device=counts.device)
    ws_bytes = lib.dmcf_scan_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=counts.device)
    fn = lib.dmcf_exclusive_scan_i32_i64 if out_dtype == torch.int64 else lib.dmcf_exclusive_scan_i32_i32
    check(fn(_p(counts), n, _p(out), _p(ws), ws_bytes, _stream()))
    return out


def neighbor_counts(points, queries, radius, ignore_query_point=False, cell_list=None):
    """Per-query neighbour count (== reduce_subarrays_sum(ones, row_splits), models/pbf_model.py:450-453)."""
    lib = _lib.load()
    nq_dev = count_of(queries)
    queries = _pos(queries, "queries")
    if cell_list is None:
        cell_list = CellList(points, max(float(radius), 1e-30))
    counts = torch.empty(queries.shape[0], dtype=torch.int32, device=queries.device)
    check(lib.dmcf_frs_count(C.byref(cell_list.grid), _p(queries), queries.shape[0], _p(nq_dev), float(radius),
                             int(bool(ignore_query_point)), _p(counts), _stream()))
    return counts, cell_list


def fixed_radius_search(points, queries, radius, ignore_query_point=False, return_distances=True, cell_list=None,
                        metric="L2", capacity=None, overflow=None):
    """CSR neighbour lists within ``radius`` (L2, inclusive).  Same outputs as the reference layer:
    neighbors_index int32 [P], neighbors_row_splits int64 [Nq+1], neighbors_distance float32 [P] (squared).

    The number of pairs is data dependent: by default it is read back from the device (ONE host sync, like the reference op's
    output allocation).  With ``capacity`` (or under a replaying StepPlan) nothing is read back: ``neighbors_index`` has
    ``capacity`` entries, the true total stays in ``neighbors_row_splits[-1]`` on the device, and a list that outgrows the
    capacity is truncated and raises ``overflow`` (int32 device tensor [1])."""
    if metric != "L2":
        raise NotImplementedError("only the L2 metric is supported (DMCF never uses another one)")
    lib = _lib.load()
    nq_dev = count_of(queries)
    points_c = _pos(points, "points")
    queries_c = _pos(queries, "queries")
    radius = float(radius)
    true_pairs = None
    if cell_list is None:  # built (and profiled) on its own
        cell_list = CellList(points, max(radius, 1e-30))
    points, queries = points_c, queries_c
    rec = _prof_begin("frs_count", n_points=points.shape[0], n_queries=queries.shape[0])
    counts = torch.empty(queries.shape[0], dtype=torch.int32, device=queries.device)
    check(lib.dmcf_frs_count(C.byref(cell_list.grid), _p(queries), queries.shape[0], _p(nq_dev), radius,
                             int(bool(ignore_query_point)), _p(counts), _stream()))
    row_splits = exclusive_scan(counts, torch.int64)
    _prof_end(rec)
    nq = queries.shape[0]
    if capacity is None and _PLAN() is not None and _PLAN().mode == "replay":
        e, slot = _PLAN().next("pairs")
        capacity = int(e["total"] * StepPlan.PAIR_SLACK) + 4096
        overflow = _PLAN().hard(slot)
        true_pairs = e["total"]  # for the profiling records only (the measured count of the planning step)
    if capacity is None:
        total = int(row_splits[-1].item())  # data-dependent output size: the one host sync of the op
        if _PLAN() is not None and _PLAN().mode == "measure":
            _PLAN().record("pairs", total=total)
    else:
        total = int(capacity)
        if overflow is not None:  # flag here as well: the fill below only sees rows, the convs only see the clamped offsets
            overflow.copy_(torch.maximum(overflow, (row_splits[-1:] > total).to(torch.int32)))
        row_splits.clamp_(max=total)  # a truncated list must never send a consumer beyond the buffers
    index = torch.empty(total, dtype=torch.int32, device=queries.device)
    dist = torch.empty(total if return_distances else 0, dtype=torch.float32, device=queries.device)
    if total > 0:
        rec = _prof_begin("frs_fill", n_points=points.shape[0], n_queries=nq, pairs=true_pairs if true_pairs is not None else total,
                          distances=bool(return_distances))
        check(lib.dmcf_frs_fill(C.byref(cell_list.grid), _p(queries), nq, _p(nq_dev), radius, int(bool(ignore_query_point)),
                                _p(row_splits), total, _p(index), _p(dist) if return_distances else None, _p(overflow),
                                _stream()))
        _prof_end(rec)
    if true_pairs is not None:
        index._dmcf_true_pairs = true_pairs
    return NeighborSearchResult(index, row_splits, dist)


def reduce_subarrays_sum_ones(row_splits):
    """reduce_subarrays_sum(ones_like(index), row_splits) as used at models/pbf_model.py:450-453."""
    return (row_splits[1:] - row_splits[:-1]).to(torch.float32)


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance,
                    neighbors_index, neighbors_importance, neighbors_row_splits, align_corners=True,
                    coordinate_mapping="ball_to_cube_radial", normalize=False, interpolation="linear",
                    max_temp_mem_MB=64, *, window=None, window_fac=1.0, relu_input=False, feat_scale=1.0,
                    ascc=False, skip_self=False, nbr_range=None, bias=None, dense_inp=None, dense_cin=0,
                    residual=None, out=None, accumulate=False, kernel_size=None, pair_records=None,
                    antisymmetric_filter=False, block_diagonal=None):
    """``ml3d.ops.continuous_conv`` (kwargs as assembled at utils/convolutions.py:414-429) plus keyword-only fused
    extras (see include/dmcf_b200.h).  ``block_diagonal=(cin_a, cout_a, cout_b)`` is the caller's promise of
    ``dmcf_conv_desc::block_cin`` (two convs over zero-padded channel groups fused into one call).  ``filters`` is [kz,ky,kx,Cin,Cout] or, with a fused Dense, the flattened
    [(kz*ky*kx*Cin + dense_cin), Cout] matrix together with ``kernel_size``."""
    lib = _lib.load()
    n_out_dev = count_of(out_positions)
    _req(filters, "filters")
    if filters.dim() == 5:
        kz, ky, kx, cin, cout = filters.shape
    else:
        if kernel_size is None or filters.dim() != 2:
            raise ValueError("flattened filters need kernel_size=[kz,ky,kx]")
        kz, ky, kx = (int(v) for v in kernel_size)
        cout = filters.shape[1]
        cin = (filters.shape[0] - dense_cin) // (kz * ky * kx)
        if cin * kz * ky * kx + dense_cin != filters.shape[0]:
            raise ValueError("flattened filter rows do not match kernel_size/dense_cin")
    filters = filters.contiguous()
    out_positions = _pos(out_positions, "out_positions")
    inp_positions = _pos(inp_positions, "inp_positions")
    inp_features, inp_stride = _rows(inp_features, "inp_features")
    if inp_features.shape[1] != cin:
        raise ValueError(f"inp_features has {inp_features.shape[1]} channels, filter expects {cin}")
    if inp_features.shape[0] != inp_positions.shape[0]:
        raise ValueError("inp_features and inp_positions disagree on the number of points")
    n_out, n_inp = out_positions.shape[0], inp_positions.shape[0]
    ext = torch.as_tensor(extents).reshape(-1)
    if ext.numel() != 1:
        raise NotImplementedError("per-point extents (RadiusSearch path) are not reachable from DMCF models")
    extent = float(ext[0])
    _req(neighbors_index, "neighbors_index", torch.int32, 1)
    _req(neighbors_row_splits, "neighbors_row_splits", torch.int64, 1)
    if neighbors_row_splits.shape[0] != n_out + 1:
        raise ValueError("neighbors_row_splits must have n_out+1 entries")
    if inp_importance is not None and inp_importance.numel() == 0:
        inp_importance = None
    if neighbors_importance is not None and neighbors_importance.numel() == 0:
        neighbors_importance = None
    if inp_importance is not None:
        _req(inp_importance, "inp_importance", dim=1)
    if neighbors_importance is not None:
        _req(neighbors_importance, "neighbors_importance", dim=1)
        if neighbors_importance.shape[0] != neighbors_index.shape[0]:
            raise ValueError("neighbors_importance and neighbors_index differ in length")
    d = ConvDesc()
    d.kernel_size[:] = [kz, ky, kx]
    d.cin, d.cout = cin, cout
    d.mapping = MAPPINGS[coordinate_mapping]
    d.interpolation = INTERPOLATIONS[interpolation]
    d.align_corners = int(bool(align_corners))
    d.normalize = int(bool(normalize))
    d.window = WINDOWS[window]
    d.window_fac = float(window_fac)
    d.extent = extent
    off = [0.0, 0.0, 0.0] if offset is None else [float(v) for v in torch.as_tensor(offset).reshape(-1).tolist()]
    d.offset[:] = off
    d.relu_input = int(bool(relu_input))
    d.feat_scale = float(feat_scale)
    d.ascc = int(bool(ascc))
    d.skip_self = int(bool(skip_self))
    d.nbr_lo, d.nbr_hi = (0, 0) if nbr_range is None else (int(nbr_range[0]), int(nbr_range[1]))
    d.dense_cin = int(dense_cin)
    d.accumulate = int(bool(accumulate))
    d.filter_antisym = int(bool(antisymmetric_filter))  # promise: filters[rev(cell)] == -filters[cell] exactly
    d.n_out_dev = None if n_out_dev is None else n_out_dev.data_ptr()
    if block_diagonal is not None:
        d.block_cin = int(block_diagonal[0])
        d.block_cout[:] = [int(block_diagonal[1]), int(block_diagonal[2])]
    dense_stride = 0
    if dense_cin:
        dense_inp, dense_stride = _rows(dense_inp, "dense_inp")
    res_stride = 0
    if residual is not None:
        residual, res_stride = _rows(residual, "residual")
    if bias is not None:
        _req(bias, "bias", dim=1)
        bias = bias.contiguous()
    if out is None:
        if accumulate:
            raise ValueError("accumulate needs an out tensor")
        out = torch.empty((n_out, cout), dtype=torch.float32, device=out_positions.device)
    _req(out, "out", dim=2)
    if out.shape[0] != n_out or out.shape[1] != cout or out.stride(1) != 1:
        raise ValueError("out has the wrong shape / layout")
    out_stride = out.stride(0) if n_out > 1 else max(cout, out.stride(0))
    neighbors_index = neighbors_index.contiguous()
    neighbors_row_splits = neighbors_row_splits.contiguous()
    if neighbors_index.numel() == 0:  # no pair at all: the ABI still wants a non-NULL list
        neighbors_index = torch.zeros(1, dtype=torch.int32, device=out_positions.device)
    rec = None
    if PROFILE is not None:
        rec = dict(kernel=conv_kernel_name((kz, ky, kx), cin, cout, interpolation, int(dense_cin), bool(antisymmetric_filter),
                                           block_diagonal, bool(ascc or normalize or nbr_range)),
                   kernel_size=(kz, ky, kx), cin=cin, cout=cout, ascc=bool(ascc), n_inp=n_inp, n_out=n_out,
                   rows=kz * ky * kx * cin + int(dense_cin),
                   pairs=int(getattr(neighbors_index, "_dmcf_true_pairs", neighbors_index.shape[0])),
                   residual=residual is not None, start=torch.cuda.Event(enable_timing=True),
                   end=torch.cuda.Event(enable_timing=True))
        rec["start"].record()
    n_pairs = 0
    if pair_records is not None:
        _req(pair_records, "pair_records", dim=2)
        n_pairs = int(neighbors_index.shape[0])
        if tuple(pair_records.shape) != (RECORD_FIELDS, n_pairs) or not pair_records.is_contiguous():
            raise ValueError(f"pair_records must be a contiguous [{RECORD_FIELDS}, n_pairs] tensor from prepare_pair_records")
    check(lib.dmcf_cconv_forward(C.byref(d), _p(filters), _p(out_positions), n_out, _p(inp_positions),
                                 _p(inp_features), inp_stride, n_inp, _p(inp_importance),
                                 _p(neighbors_index), _p(neighbors_row_splits),
                                 _p(neighbors_importance), _p(bias), _p(dense_inp) if dense_cin else None,
                                 dense_stride, _p(residual), res_stride, _p(out), out_stride, _p(pair_records), n_pairs,
                                 _stream()))
    if rec is not None:
        rec["end"].record()
        PROFILE.append(rec)
    return out


def conv_patches(kernel_size, out_positions, extents, inp_positions, inp_features, neighbors_index, neighbors_row_splits,
                 align_corners=True, coordinate_mapping="ball_to_cube_radial", interpolation="linear", *, window=None,
                 window_fac=1.0, relu_input=False, feat_scale=1.0, skip_self=False, neighbors_importance=None, out=None):
    """The patch matrix of a conv (dmcf_cconv_patches): ``[n_out, kz*ky*kx*cin]`` with
    ``continuous_conv(filters, ...) == patches @ filters.reshape(-1, cout)``; the gradient of the conv w.r.t. its filter
    is ``patches.T @ d_out``.  ``out_positions`` / ``neighbors_row_splits`` may be a slice [a:b] / [a:b+1] of the full
    arrays (row_splits hold absolute offsets into ``neighbors_index``)."""
    lib = _lib.load()
    out_positions = _pos(out_positions, "out_positions")
    inp_positions = _pos(inp_positions, "inp_positions")
    inp_features, inp_stride = _rows(inp_features, "inp_features")
    _req(neighbors_index, "neighbors_index", torch.int32, 1)
    _req(neighbors_row_splits, "neighbors_row_splits", torch.int64, 1)
    neighbors_index = neighbors_index.contiguous()
    neighbors_row_splits = neighbors_row_splits.contiguous()
    n_out, n_inp, cin = out_positions.shape[0], inp_positions.shape[0], inp_features.shape[1]
    if neighbors_row_splits.shape[0] != n_out + 1:
        raise ValueError("neighbors_row_splits must have n_out+1 entries")
    kz, ky, kx = (int(k) for k in kernel_size)
    d = ConvDesc()
    d.kernel_size[:] = [kz, ky, kx]
    d.cin, d.cout = cin, 4
    d.mapping = MAPPINGS[coordinate_mapping]
    d.interpolation = INTERPOLATIONS[interpolation]
    d.align_corners = int(bool(align_corners))
    d.normalize = 0
    d.window = WINDOWS[window]
    d.window_fac = float(window_fac)
    d.extent = float(torch.as_tensor(extents).reshape(-1)[0])
    d.offset[:] = [0.0, 0.0, 0.0]
    d.relu_input = int(bool(relu_input))
    d.feat_scale = float(feat_scale)
    d.skip_self = int(bool(skip_self))
    if neighbors_importance is not None:
        _req(neighbors_importance, "neighbors_importance", dim=1)
        neighbors_importance = neighbors_importance.contiguous()
    kc = kz * ky * kx * cin
    if out is None:
        out = torch.empty((n_out, kc), dtype=torch.float32, device=out_positions.device)
    _req(out, "out", dim=2)
    if out.shape[0] != n_out or out.shape[1] != kc or out.stride(1) != 1:
        raise ValueError("out has the wrong shape / layout")
    if neighbors_index.numel() == 0:
        neighbors_index = torch.zeros(1, dtype=torch.int32, device=out_positions.device)
    check(lib.dmcf_cconv_patches(C.byref(d), _p(out_positions), n_out, _p(inp_positions), _p(inp_features), inp_stride,
                                 n_inp, None, _p(neighbors_index), _p(neighbors_row_splits), _p(neighbors_importance), None,
                                 0, _p(out), out.stride(0) if n_out > 1 else kc, _stream()))
    return out


def prepare_pair_records(kernel_size, out_positions, extents, offset, inp_positions, inp_importance, neighbors_index,
                         neighbors_importance, neighbors_row_splits, align_corners=True,
                         coordinate_mapping="ball_to_cube_radial", interpolation="linear", *, window=None, window_fac=1.0,
                         skip_self=False, nbr_range=None):
    """Per-pair geometry of a neighbour list, evaluated once and shared by every conv with the same geometry arguments
    (dmcf_cconv_prepare).  Returns a [9, n_pairs] float32 tensor to pass as ``pair_records`` to continuous_conv."""
    lib = _lib.load()
    n_out_dev = count_of(out_positions)
    out_positions = _pos(out_positions, "out_positions")
    inp_positions = _pos(inp_positions, "inp_positions")
    _req(neighbors_index, "neighbors_index", torch.int32, 1)
    _req(neighbors_row_splits, "neighbors_row_splits", torch.int64, 1)
    neighbors_index = neighbors_index.contiguous()
    neighbors_row_splits = neighbors_row_splits.contiguous()
    if inp_importance is not None and inp_importance.numel() == 0:
        inp_importance = None
    if neighbors_importance is not None and neighbors_importance.numel() == 0:
        neighbors_importance = None
    d = ConvDesc()
    d.kernel_size[:] = [int(k) for k in kernel_size]
    d.cin, d.cout = 1, 1
    d.mapping = MAPPINGS[coordinate_mapping]
    d.interpolation = INTERPOLATIONS[interpolation]
    d.align_corners = int(bool(align_corners))
    d.window = WINDOWS[window]
    d.window_fac = float(window_fac)
    d.extent = float(torch.as_tensor(extents).reshape(-1)[0])
    d.offset[:] = [0.0, 0.0, 0.0] if offset is None else [float(v) for v in torch.as_tensor(offset).reshape(-1).tolist()]
    d.feat_scale = 1.0
    d.skip_self = int(bool(skip_self))
    d.nbr_lo, d.nbr_hi = (0, 0) if nbr_range is None else (int(nbr_range[0]), int(nbr_range[1]))
    d.n_out_dev = None if n_out_dev is None else n_out_dev.data_ptr()
    n_pairs = int(neighbors_index.shape[0])
    records = torch.empty((RECORD_FIELDS, n_pairs), dtype=torch.float32, device=out_positions.device)
    rec = _prof_begin("pair_records", n_inp=inp_positions.shape[0], n_out=out_positions.shape[0],
                      pairs=int(getattr(neighbors_index, "_dmcf_true_pairs", n_pairs)))
    check(lib.dmcf_cconv_prepare(C.byref(d), _p(out_positions), out_positions.shape[0], _p(inp_positions),
                                 inp_positions.shape[0], _p(inp_importance), _p(neighbors_index), _p(neighbors_row_splits),
                                 _p(neighbors_importance), n_pairs, _p(records), _stream()))
    _prof_end(rec)
    return records


def dense(x, kernel, bias=None, relu_input=False, out=None):
    """Keras Dense (linear): g(x) @ kernel[Cin,Cout] + bias."""
    lib = _lib.load()
    x, xs = _rows(x, "x")
    _req(kernel, "kernel", dim=2)
    kernel = kernel.contiguous()
    cin, cout = kernel.shape
    if x.shape[1] != cin:
        raise ValueError("x / kernel channel mismatch")
    n = x.shape[0]
    if out is None:
        out = torch.empty((n, cout), dtype=torch.float32, device=x.device)
    _req(out, "out", dim=2)
    if out.shape[0] != n or out.shape[1] != cout or out.stride(1) != 1:
        raise ValueError("out has the wrong shape / layout")
    out_stride = out.stride(0) if n > 1 else max(cout, out.stride(0))
    if bias is not None:
        bias = _req(bias, "bias", dim=1).contiguous()
    check(lib.dmcf_dense_forward(_p(x), n, cin, xs, _p(kernel), _p(bias), cout, int(bool(relu_input)), _p(out),
                                 out_stride, _stream()))
    return out


def integrate(pos, vel, acc, gravity, dt):
    """models/pbf_model.py:234-240."""
    lib = _lib.load()
    pos, vel = _pos(pos, "pos"), _pos(vel, "vel")
    acc = None if acc is None else _pos(acc, "acc")
    g = (C.c_float * 3)(*[float(v) for v in gravity])
    pos2, vel2 = torch.empty_like(pos), torch.empty_like(vel)
    check(lib.dmcf_integrate(_p(pos), _p(vel), _p(acc), g, float(dt), pos.shape[0], _p(pos2), _p(vel2), _stream()))
    return pos2, vel2


def correct(pos, pos2, net, out_scale, dt):
    """models/pbf_model.py:466-487 (channel expansion, out_scale, position correction, velocity update)."""
    lib = _lib.load()
    pos, pos2 = _pos(pos, "pos"), _pos(pos2, "pos2")
    net, ns = _rows(net, "net")
    if net.shape[0] < pos.shape[0]:
        raise ValueError("net has fewer rows than particles")
    s = (C.c_float * 3)(*[float(v) for v in out_scale])
    pos_new, vel_new = torch.empty_like(pos), torch.empty_like(pos)
    check(lib.dmcf_correct(_p(pos), _p(pos2), _p(net), ns, net.shape[1], s, float(dt), pos.shape[0], _p(pos_new),
                           _p(vel_new), _stream()))
    return pos_new, vel_new


def grid_pos(pos, voxel, center=None, hyst=0.1):
    """Lattice sampling of utils/tools/losses.py:136-181 (``center`` is None when not centralised).

    The number of lattice points (and the integer bounds of the lattice) are data dependent: by default they are read back from
    the device (two host syncs).  Under a replaying StepPlan the bounds and the capacity come from the plan, ``center`` stays on
    the device and the result is a capacity-sized tensor with its count attached (``count_of``)."""
    import numpy as np
    lib = _lib.load()
    n_dev = count_of(pos)
    pos = _pos(pos, "pos")
    n = pos.shape[0]
    f32 = np.float32
    v = np.asarray(voxel, f32).reshape(3)
    cv = (C.c_float * 3)(*[float(x) for x in v])
    if _PLAN() is not None and _PLAN().mode == "replay":
        e, slot = _PLAN().next("lattice")
        active = (v >= f32(1e-5)).astype(np.int64)
        pad = (4 + np.asarray(e["dims"], np.int64) // 10) * active  # voxels of slack on every side of the measured lattice
        lo = np.asarray(e["lo"], np.int64) - pad
        dims = np.asarray(e["dims"], np.int64) + 2 * pad
        capacity = int(e["count"] * StepPlan.ROW_SLACK) + 64
        n_cells = int(dims[0]) * int(dims[1]) * int(dims[2])
        clo = (C.c_int32 * 3)(*[int(x) for x in lo])
        cd = (C.c_int32 * 3)(*[int(x) for x in dims])
        c_dev = None if center is None else _req(center, "center", dim=1).contiguous()
        flags = torch.zeros(n_cells, dtype=torch.int32, device=pos.device)
        ovf = _PLAN().hard(slot)
        check(lib.dmcf_grid_pos_mark(_p(pos), n, _p(n_dev), cv, None, _p(c_dev), float(hyst), clo, cd, _p(flags), _p(ovf), _stream()))
        offsets = exclusive_scan(flags, torch.int32)
        out = torch.zeros((capacity, 3), dtype=torch.float32, device=pos.device)
        check(lib.dmcf_grid_pos_emit(_p(flags), _p(offsets), cv, None, _p(c_dev), clo, cd, _p(out), capacity, _p(ovf), _stream()))
        return with_count(out, torch.clamp(offsets[-1:], max=capacity))
    if n_dev is not None:
        raise DmcfError("grid_pos of a capacity-sized point set needs a replaying StepPlan")
    if n == 0:
        if _PLAN() is not None and _PLAN().mode == "measure":
            raise DmcfError("cannot plan a step on an empty particle set")
        return torch.empty((0, 3), dtype=torch.float32, device=pos.device)
    stats = [pos.amin(dim=0), pos.amax(dim=0)]
    if center is not None:
        stats.append(_req(center, "center", dim=1))
    stats = torch.stack(stats).cpu().numpy().astype(f32)  # one host sync (output size is data dependent anyway)
    c = stats[2] if center is not None else np.zeros(3, f32)
    vm = np.maximum(v, f32(1e-5))
    h = np.where(v >= f32(1e-5), f32(hyst), f32(0)).astype(f32)
    active = (v >= f32(1e-5)).astype(np.int64)
    with np.errstate(over="ignore"):
        lo = np.floor(((stats[0] - c).astype(f32) / vm).astype(f32) - h).astype(np.int64)
        hi = np.floor(((stats[1] - c).astype(f32) / vm).astype(f32) + h).astype(np.int64) + active
    dims = hi - lo + 1
    n_cells = int(dims[0]) * int(dims[1]) * int(dims[2])
    if n_cells >= 2 ** 31 or np.any(np.abs(lo) >= 2 ** 30):
        raise DmcfError(f"grid_pos lattice of {dims.tolist()} voxels is too large")
    cc = (C.c_float * 3)(*[float(x) for x in c]) if center is not None else None
    clo = (C.c_int32 * 3)(*[int(x) for x in lo])
    cd = (C.c_int32 * 3)(*[int(x) for x in dims])
    flags = torch.zeros(n_cells, dtype=torch.int32, device=pos.device)
    check(lib.dmcf_grid_pos_mark(_p(pos), n, None, cv, cc, None, float(hyst), clo, cd, _p(flags), None, _stream()))
    offsets = exclusive_scan(flags, torch.int32)
    count = int(offsets[-1].item())
    if _PLAN() is not None and _PLAN().mode == "measure":
        _PLAN().record("lattice", lo=[int(x) for x in lo], dims=[int(x) for x in dims], count=count)
    out = torch.empty((count, 3), dtype=torch.float32, device=pos.device)
    if count:
        check(lib.dmcf_grid_pos_emit(_p(flags), _p(offsets), cv, cc, None, clo, cd, _p(out), -1, None, _stream()))
    return out


def set_kernel_options(options):
    """bit 0: register-patch kernels for wide layers (k_cconv_ws / k_cconv_lean, default on), bit 1: direct kernel for cout <= 4,
    bit 2: z-split launches of the legacy k_cconv_wide, bit 3: legacy k_cconv_wide instead of k_cconv_lean, bit 5: single-pair
    walk for narrow inputs, bit 6: query-centric search for prefix searches, bit 7: the warp-specialised k_cconv_ws instead of
    k_cconv_lean (a measured experiment, slower), bit 12: no narrow direct kernel (k_cconv_narrow), bit 13: SIMT Dense, bit 14: one CTA per tile in the tensor-core
    k_cconv_lean, bit 15: FFMA2 instead of its tensor-core phase 2, bit 16: its 16-point / 16-warp tile.  Returns the previous mask."""
    return int(_lib.load().dmcf_set_kernel_options(int(options)))


RECORD_FIELDS = 9  # dmcf_cconv_prepare: {row, i0, i1, wx0, wx1, wy0, wy1, wz0*a, wz1*a} per pair


def conv_kernel_name(kernel_size, cin, cout, interpolation, dense_cin=0, antisymmetric_filter=False, block_diagonal=None,
                     not_narrow=False):
    """Which kernel dmcf_cconv_forward dispatches to with the default options (mirrors csrc/cconv.cu)."""
    kz, ky, kx = (int(k) for k in kernel_size)
    kc = kz * ky * kx * cin + dense_cin
    if (block_diagonal is not None and not not_narrow and cout <= 32 and dense_cin <= 32 and kz * ky * kx < 512
            and block_diagonal[0] <= 4 and 1 <= cin - block_diagonal[0] <= 4 and max(block_diagonal[1:]) <= 8):
        return "k_cconv_narrow"
    if (antisymmetric_filter and interpolation == "linear" and cin <= 32
            and ((kz, ky, kx), cout) == ((1, 8, 8), 2)):
        return "k_cconv_apatch"
    if (cout <= 4 and interpolation == "linear" and cin <= 32 and (kz, ky, kx) in ((4, 4, 4), (1, 8, 8), (1, 8, 1))):
        return "k_cconv_apatch"
    if cout <= 4 and (kc * cout + 4 + 16 * 32 * 12) * 4 <= 200 * 1024:
        return "k_cconv_direct"
    if interpolation == "linear" and cin <= 32 and cout % 4 == 0 and (kz, ky, kx) in ((4, 4, 4), (1, 8, 8), (1, 8, 1)):
        kc_pad = (kc + 3) // 4 * 4
        lean_smem = (max(kc_pad // 4 * 25 * 4, 12 * 24 * 32) + 12 * 384 + 24) * 4
        return "k_cconv_lean" if cout <= 32 and lean_smem <= 227 * 1024 else "k_cconv_wide"
    return "k_cconv_tile"


def launch_count():
    return int(_lib.load().dmcf_launch_count())
