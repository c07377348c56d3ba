"""Spatial slab decomposition across GPUs with one-hop halo exchange.  The reference is single device
(run_pipeline.py:84-100); this is the multi-GPU design the north star asks for.

One process per GPU.  Rank k owns the particles whose coordinate along ``axis`` lies in [faces[k], faces[k+1]).
Per step:
  * bbox / mean all-reduces (the boundary cull of models/pbf_model.py:330-336 and ``centralize`` need global values),
  * one POSITION halo: every rank sends the owned points within ``radius`` of a face to the rank behind that face,
  * one FEATURE halo per conv layer: the same rows of the layer's input features (lists fixed for the step),
  * MIGRATION of particles that crossed a face after the position update.
Convs then run on [owned | ghosts] inputs and owned outputs; pair terms are evaluated from bit-identical operands on
both sides of a face, so the antisymmetric layer conserves momentum across ranks exactly as on one GPU.

Everything here is plain torch + torch.distributed point-to-point (NCCL on GPUs, gloo in the CPU tests); payloads
are a few MB per face per layer, i.e. latency bound, so they are batched into one isend/irecv group per exchange.
"""
from __future__ import annotations

import queue
import threading

import torch
import torch.distributed as dist


class DistTransport:
    """The production transport: torch.distributed point-to-point / all-reduce (NCCL on GPUs, gloo in the CPU tests)."""

    graph_capturable = True  # NCCL send / recv / all-reduce can be captured in a CUDA graph (gloo cannot)

    def __init__(self, group=None):
        self.group = group
        if dist.is_initialized() and dist.get_backend(group) != "nccl":
            self.graph_capturable = False

    def all_reduce(self, rank, t, op):
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.group)
        return t

    def sendrecv(self, rank, sends, recvs):
        """``sends`` / ``recvs``: lists of (peer rank, tensor) in matching order on both sides; receives are filled in place.
        One batched isend/irecv group; waiting on it orders the current stream behind the transfer (no host sync with NCCL)."""
        ops = []
        for peer, t in sends:
            if t.numel():
                ops.append(dist.P2POp(dist.isend, t, peer, self.group))
        for peer, t in recvs:
            if t.numel():
                ops.append(dist.P2POp(dist.irecv, t, peer, self.group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class LocalTransport:
    """In-process stand-in for torch.distributed: ``world`` SlabContexts, each driven by its own Python thread on ONE device
    (the default stream orders the work of all threads, and a tensor is handed over only after its producer kernels were
    enqueued).  Lets the 2-slab == 1-slab parity test run on a single GPU (tests/test_slab_gpu.py) and the halo / migration
    logic on CPU tensors."""

    def __init__(self, world):
        self.world = world
        self._barrier = threading.Barrier(world)
        self._slots = [None] * world
        self._q = {(a, b): queue.Queue() for a in range(world) for b in range(world) if a != b}

    def all_reduce(self, rank, t, op):
        self._slots[rank] = t.clone()
        self._barrier.wait()
        parts = torch.stack(list(self._slots))
        res = parts.amax(dim=0) if op == "max" else parts.sum(dim=0)
        self._barrier.wait()  # every rank has read the slots before anyone overwrites them
        t.copy_(res)
        return t

    def sendrecv(self, rank, sends, recvs):
        for peer, t in sends:
            self._q[(rank, peer)].put(t.clone())
        for peer, t in recvs:
            src = self._q[(peer, rank)].get(timeout=120)
            if src.shape != t.shape:
                raise RuntimeError(f"LocalTransport: rank {rank} expected {tuple(t.shape)} from {peer}, got {tuple(src.shape)}")
            t.copy_(src)


def _ops():
    from . import ops  # CUDA-only module: imported lazily so that the halo logic also runs on CPU tensors (tests/test_slab_cpu.py)
    return ops


def _plan():
    import sys
    mod = sys.modules.get("dmcf_b200.ops")
    return mod.get_plan() if mod is not None else None


def _count_of(t):
    return getattr(t, "_dmcf_n_dev", None)


def _valid(t):
    """Bool mask of the valid rows of a capacity-sized tensor, or None when every row is valid."""
    cnt = _count_of(t)
    if cnt is None:
        return None
    return torch.arange(t.shape[0], device=t.device, dtype=torch.int32) < cnt


class SlabContext:
    def __init__(self, faces, axis=0, group=None, rank=None, world_size=None, transport=None):
        """``faces``: world_size+1 increasing coordinates (use -inf / +inf for the outer ones).  ``transport``: a
        LocalTransport shared by the ranks of a single-process run (then ``rank`` / ``world_size`` are required)."""
        self.group = group
        self.transport = transport if transport is not None else DistTransport(group)
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world_size is None else world_size
        if len(faces) != self.world + 1:
            raise ValueError("need world_size+1 slab faces")
        self.faces = [float(f) for f in faces]
        self.axis = axis
        self.lo, self.hi = self.faces[self.rank], self.faces[self.rank + 1]
        self.left = self.rank - 1 if self.rank > 0 else None
        self.right = self.rank + 1 if self.rank < self.world - 1 else None
        self._plan = None
        self.bytes_exchanged = 0

    # -- ownership -------------------------------------------------------------------------------------------
    @staticmethod
    def uniform_faces(lo, hi, world):
        inf = float("inf")
        inner = [lo + (hi - lo) * k / world for k in range(1, world)]
        return [-inf] + inner + [inf]

    def owner_of(self, pos):
        """Rank owning each position (bucketize on the inner faces)."""
        inner = torch.tensor(self.faces[1:-1], dtype=pos.dtype, device=pos.device)
        return torch.bucketize(pos[:, self.axis].contiguous(), inner, right=True)

    def owned_mask(self, pos):
        x = pos[:, self.axis]
        m = (x >= self.lo) & (x < self.hi)
        v = _valid(pos)
        return m if v is None else (m & v)

    # -- collectives -----------------------------------------------------------------------------------------
    def all_reduce_minmax(self, lo, hi):
        """Global bounding box of per-rank (lo, hi) 3-vectors."""
        if self.world == 1:
            return lo, hi
        buf = torch.cat([-lo, hi]).contiguous()
        self.transport.all_reduce(self.rank, buf, "max")
        return -buf[:3], buf[3:]

    def all_reduce_sum(self, t):
        if self.world > 1:
            self.transport.all_reduce(self.rank, t, "sum")
        return t

    def all_reduce_max(self, t):
        if self.world > 1:
            self.transport.all_reduce(self.rank, t, "max")
        return t

    def _exchange(self, to_left, to_right, known=None):
        """Sends row blocks to the two neighbours, returns (from_left, from_right, known).

        Eager / measuring step: the row counts are exchanged first and read back (one host sync) unless ``known`` (the
        counts of an earlier exchange of the same rows: feature halos reuse those of the position halo).  Replaying step
        (dmcf_b200.ops.StepPlan): both sides derive the same message capacity from the count measured when the plan was made,
        the true counts travel as device tensors next to the payload, nothing is read back."""
        dev, dt = to_left.device, to_left.dtype
        cols = to_left.shape[1:]
        plan = _plan()
        peers = [p for p in (self.left, self.right) if p is not None]
        payload = {self.left: to_left.contiguous(), self.right: to_right.contiguous()}
        if plan is not None and plan.mode == "replay":
            ops = _ops()
            if known is None:
                caps = [ops.planned_rows() for _ in (0, 1)]  # (capacity, overflow flag) of the arrivals from left / right
                cnt_in = {p: torch.zeros(1, dtype=torch.int32, device=dev) for p in peers}
                cnt_out = {p: _count_of(payload[p]) for p in peers}
                self.transport.sendrecv(self.rank, [(p, cnt_out[p]) for p in peers], [(p, cnt_in[p]) for p in peers])
                known = dict(cap_left=caps[0][0], cap_right=caps[1][0], cnt_left=cnt_in.get(self.left), cnt_right=cnt_in.get(self.right))
            from_left = torch.zeros((known["cap_left"] if self.left is not None else 0, *cols), dtype=dt, device=dev)
            from_right = torch.zeros((known["cap_right"] if self.right is not None else 0, *cols), dtype=dt, device=dev)
            recv = {self.left: from_left, self.right: from_right}
            self.transport.sendrecv(self.rank, [(p, payload[p]) for p in peers], [(p, recv[p]) for p in peers])
            zero = torch.zeros(1, dtype=torch.int32, device=dev)
            from_left = ops.with_count(from_left, known["cnt_left"] if self.left is not None else zero)
            from_right = ops.with_count(from_right, known["cnt_right"] if self.right is not None else zero)
            self.bytes_exchanged += sum(payload[p].numel() for p in peers) * to_left.element_size()
            return from_left, from_right, known
        if known is None:
            cnt_out = torch.tensor([to_left.shape[0], to_right.shape[0]], dtype=torch.int64, device=dev)
            cnt_l = torch.zeros(1, dtype=torch.int64, device=dev)
            cnt_r = torch.zeros(1, dtype=torch.int64, device=dev)
            sends, recvs = [], []
            if self.left is not None:
                sends.append((self.left, cnt_out[0:1]))
                recvs.append((self.left, cnt_l))
            if self.right is not None:
                sends.append((self.right, cnt_out[1:2]))
                recvs.append((self.right, cnt_r))
            self.transport.sendrecv(self.rank, sends, recvs)
            known = dict(n_left=int(cnt_l.item()), n_right=int(cnt_r.item()))
            if plan is not None:  # measuring step: the arrivals' counts are sizes of the plan
                ops = _ops()
                ops.planned_rows(known["n_left"])
                ops.planned_rows(known["n_right"])
        from_left = torch.empty((known["n_left"], *cols), dtype=dt, device=dev)
        from_right = torch.empty((known["n_right"], *cols), dtype=dt, device=dev)
        recv = {self.left: from_left, self.right: from_right}
        self.transport.sendrecv(self.rank, [(p, payload[p]) for p in peers], [(p, recv[p]) for p in peers])
        self.bytes_exchanged += sum(payload[p].numel() for p in peers) * to_left.element_size()
        return from_left, from_right, known

    def _select(self, mask, *tensors, bounded=True):
        """Rows where ``mask`` holds: boolean indexing, or (on CUDA under a StepPlan) the plan-aware capacity version."""
        if _plan() is not None:
            return _ops().select_rows(mask, *tensors, bounded=bounded)
        return [t[mask] for t in tensors]

    # -- halos -----------------------------------------------------------------------------------------------
    def make_halo(self, pos, radius):
        """Halo plan of one point set (particles, or the lattice points of one scale) for this step: which owned rows
        go to which neighbour, how many arrive, and the ghost positions [from left ; from right].  A point within
        ``radius`` (padded like the search) of a face is sent."""
        x = pos[:, self.axis]
        pad = float(radius) * 1.0001 + 1e-6 * max(abs(self.lo) if self.lo > -1e30 else 0.0, abs(self.hi) if self.hi < 1e30 else 0.0)
        valid = _valid(pos)
        none = torch.zeros(pos.shape[0], dtype=torch.bool, device=pos.device)
        m_left = (x < self.lo + pad) if self.left is not None else none
        m_right = (x >= self.hi - pad) if self.right is not None else none
        if valid is not None:
            m_left, m_right = m_left & valid, m_right & valid
        rows = torch.arange(pos.shape[0], device=pos.device, dtype=torch.int64)
        plan = HaloPlan(self)
        (plan.send_left,) = self._select(m_left, rows, bounded=False)
        (plan.send_right,) = self._select(m_right, rows, bounded=False)
        gl, gr, plan.known = self._exchange(self._gather(pos, plan.send_left), self._gather(pos, plan.send_right))
        plan.ghost_left, plan.ghost_right = gl, gr
        plan.ghost_pos = self._cat([gl, gr])
        return plan

    @staticmethod
    def _gather(t, idx):
        out = t[idx]
        cnt = _count_of(idx)
        if cnt is not None:
            out._dmcf_n_dev = cnt
        return out

    @staticmethod
    def _cat(parts):
        if all(_count_of(t) is None for t in parts):
            return torch.cat(parts, dim=0)
        return _ops().concat_rows(parts)

    # single-plan convenience API (one point set per step)
    def position_halo(self, pos, radius):
        self._plan = self.make_halo(pos, radius)
        return self._plan.ghost_pos

    def feature_halo(self, feats):
        return self._plan.feature_halo(feats)

    def with_ghosts(self, feats):
        if self.world == 1:
            return feats
        return self._plan.with_ghosts(feats)

    # -- migration -------------------------------------------------------------------------------------------
    def migrate(self, pos, *others):
        """Re-establishes ownership after a position update: rows that left the slab go to the neighbour (one hop per
        step: speed*dt << slab width), arrivals are appended.  Returns the new (pos, *others)."""
        if self.world == 1:
            return (pos, *others)
        x = pos[:, self.axis]
        valid = _valid(pos)
        go_l = x < self.lo
        go_r = x >= self.hi
        if self.left is None:
            go_l = torch.zeros_like(go_l)
        if self.right is None:
            go_r = torch.zeros_like(go_r)
        stay = ~(go_l | go_r)
        if valid is not None:
            go_l, go_r, stay = go_l & valid, go_r & valid, stay & valid
        packed = torch.cat([pos] + [o.reshape(o.shape[0], -1) for o in others], dim=1)
        (out_l,) = self._select(go_l, packed, bounded=False)
        (out_r,) = self._select(go_r, packed, bounded=False)
        fl, fr, _ = self._exchange(out_l, out_r)
        # (unbounded: the capacity of the kept rows -- and with it the shape of the returned state -- depends on the plan only,
        # not on how many rows the input buffer happened to have: a graph captured for that shape keeps fitting)
        (kept,) = self._select(stay, packed, bounded=False)
        new = self._cat([kept, fl, fr])
        cnt = _count_of(new)
        outs, c = [], 0
        for t in (pos, *others):
            w = t.reshape(t.shape[0], -1).shape[1]
            part = new[:, c:c + w].reshape(-1, *t.shape[1:])
            if cnt is not None:
                part = part.contiguous()
                part._dmcf_n_dev = cnt
            outs.append(part)
            c += w
        return tuple(outs)


class HaloPlan:
    """Row lists / counts of one point set's halo, fixed for a step; refreshes ghost FEATURE rows on demand."""

    def __init__(self, slab):
        self.slab = slab
        self.send_left = self.send_right = None
        self.known = None
        self.ghost_left = self.ghost_right = None
        self.ghost_pos = None

    def feature_halo(self, feats):
        """Ghost rows of ``feats`` (same order as ``ghost_pos``)."""
        gl, gr, _ = self.slab._exchange(SlabContext._gather(feats, self.send_left), SlabContext._gather(feats, self.send_right),
                                        self.known)
        return SlabContext._cat([gl, gr])

    def with_ghosts(self, feats):
        if self.slab.world == 1:
            return feats
        return SlabContext._cat([feats, self.feature_halo(feats)])

    def fill_ghosts(self, feats, n_own):
        """In-place variant for a feature buffer that has room behind its owned rows (rows = owned capacity + ghost
        capacities, like the [owned | ghost] position array of the same point set): the ghost rows of this layer's input go to
        rows [n_own, n_own + ghosts).  ``n_own``: int (exact) or int32 device tensor [1]."""
        gl, gr, _ = self.slab._exchange(SlabContext._gather(feats, self.send_left), SlabContext._gather(feats, self.send_right),
                                        self.known)
        if isinstance(n_own, torch.Tensor) or _count_of(gl) is not None or _count_of(gr) is not None:
            ops = _ops()
            base = ops.append_rows(feats, n_own, gl)
            ops.append_rows(feats, base, gr, want_count=False)
        else:
            feats[n_own:n_own + gl.shape[0]] = gl
            feats[n_own + gl.shape[0]:n_own + gl.shape[0] + gr.shape[0]] = gr
        return feats
