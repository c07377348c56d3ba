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

    def __init__(self, group=None):
        self.group = group

    def all_reduce(self, rank, t, op):
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=self.group)
        return t

    def sendrecv(self, rank, sends, recvs):
        """``sends`` / ``recvs``: {peer rank: tensor}; receives are filled in place.  One batched isend/irecv group."""
        ops = []
        for peer, t in sends.items():
            if t.numel():
                ops.append(dist.P2POp(dist.isend, t, peer, self.group))
        for peer, t in recvs.items():
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
        for peer, t in sends.items():
            self._q[(rank, peer)].put(t.clone())
        for peer, t in recvs.items():
            src = self._q[(peer, rank)].get(timeout=120)
            if src.shape != t.shape:
                raise RuntimeError(f"LocalTransport: rank {rank} expected {tuple(t.shape)} from {peer}, got {tuple(src.shape)}")
            t.copy_(src)


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
        return (x >= self.lo) & (x < self.hi)

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

    def _exchange(self, to_left, to_right, n_from_left=None, n_from_right=None):
        """Sends row blocks to the two neighbours, returns (from_left, from_right).  Row counts are exchanged first
        unless the caller already knows them (feature halos reuse the counts of the position halo)."""
        dev, dt = to_left.device, to_left.dtype
        cols = to_left.shape[1:]
        if n_from_left is None:
            cnt_out = torch.tensor([to_left.shape[0], to_right.shape[0]], dtype=torch.int64, device=dev)
            cnt_l = torch.zeros(1, dtype=torch.int64, device=dev)
            cnt_r = torch.zeros(1, dtype=torch.int64, device=dev)
            sends, recvs = {}, {}
            if self.left is not None:
                sends[self.left], recvs[self.left] = cnt_out[0:1], cnt_l
            if self.right is not None:
                sends[self.right], recvs[self.right] = cnt_out[1:2], cnt_r
            self.transport.sendrecv(self.rank, sends, recvs)
            n_from_left, n_from_right = int(cnt_l.item()), int(cnt_r.item())
        from_left = torch.empty((n_from_left, *cols), dtype=dt, device=dev)
        from_right = torch.empty((n_from_right, *cols), dtype=dt, device=dev)
        to_left, to_right = to_left.contiguous(), to_right.contiguous()
        sends, recvs = {}, {}
        if self.left is not None:
            sends[self.left], recvs[self.left] = to_left, from_left
        if self.right is not None:
            sends[self.right], recvs[self.right] = to_right, from_right
        self.transport.sendrecv(self.rank, sends, recvs)
        self.bytes_exchanged += (to_left.numel() + to_right.numel()) * to_left.element_size()
        return from_left, from_right, n_from_left, n_from_right

    # -- halos -----------------------------------------------------------------------------------------------
    def make_halo(self, pos, radius):
        """Halo plan of one point set (particles, or the lattice points of one scale) for this step: which owned rows
        go to which neighbour, how many arrive, and the ghost positions [from left ; from right].  A point within
        ``radius`` (padded like the search) of a face is sent."""
        x = pos[:, self.axis]
        pad = float(radius) * 1.0001 + 1e-6 * max(abs(self.lo) if self.lo > -1e30 else 0.0, abs(self.hi) if self.hi < 1e30 else 0.0)
        empty = torch.zeros(0, dtype=torch.int64, device=pos.device)
        plan = HaloPlan(self)
        plan.send_left = torch.nonzero(x < self.lo + pad).flatten() if self.left is not None else empty
        plan.send_right = torch.nonzero(x >= self.hi - pad).flatten() if self.right is not None else empty
        gl, gr, plan.n_from_left, plan.n_from_right = self._exchange(pos[plan.send_left], pos[plan.send_right])
        plan.ghost_pos = torch.cat([gl, gr], dim=0)
        return plan

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
        go_l = x < self.lo
        go_r = x >= self.hi
        stay = ~(go_l | go_r)
        packed = torch.cat([pos] + [o.reshape(o.shape[0], -1) for o in others], dim=1)
        fl, fr, _, _ = self._exchange(packed[go_l], packed[go_r])
        new = torch.cat([packed[stay], fl, fr], dim=0)
        outs, c = [], 0
        for t in (pos, *others):
            w = t.reshape(t.shape[0], -1).shape[1]
            outs.append(new[:, c:c + w].reshape(-1, *t.shape[1:]))
            c += w
        return tuple(outs)


class HaloPlan:
    """Row lists / counts of one point set's halo, fixed for a step; refreshes ghost FEATURE rows on demand."""

    def __init__(self, slab):
        self.slab = slab
        self.send_left = self.send_right = None
        self.n_from_left = self.n_from_right = 0
        self.ghost_pos = None

    def feature_halo(self, feats):
        """Ghost rows of ``feats`` (same order as ``ghost_pos``)."""
        gl, gr, _, _ = self.slab._exchange(feats[self.send_left], feats[self.send_right], self.n_from_left, self.n_from_right)
        return torch.cat([gl, gr], dim=0)

    def with_ghosts(self, feats):
        if self.slab.world == 1:
            return feats
        return torch.cat([feats, self.feature_halo(feats)], dim=0)
