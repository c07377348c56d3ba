"""Seeded synthetic scenes: the workloads bench.py times and the parity tests check.  NumPy only."""
from __future__ import annotations

import numpy as np


def _walls(lo, hi, dx, dims, open_top=False, gap=None):
    """One layer of wall particles (pitch dx) ``gap`` (default half a pitch) outside the axis-aligned block [lo,hi], inward
    normals.  Only the axes listed in ``dims`` get walls (2-D scenes live in the xy-plane, 1-D ones on y)."""
    pts, nrm = [], []
    axes = list(dims)
    gap = dx / 2 if gap is None else gap
    rng_ax = {a: np.arange(lo[a] - gap, hi[a] + gap + dx / 2, dx) for a in axes}
    for a in axes:
        for side, coord in ((+1, lo[a] - gap), (-1, hi[a] + gap)):
            if open_top and a == 1 and side == -1:
                continue
            others = [b for b in axes if b != a]
            grids = np.meshgrid(*[rng_ax[b] for b in others], indexing="ij") if others else []
            n = grids[0].size if others else 1
            p = np.zeros((n, 3))
            p[:, a] = coord
            for b, g in zip(others, grids):
                p[:, b] = g.reshape(-1)
            q = np.zeros((n, 3))
            q[:, a] = side
            pts.append(p)
            nrm.append(q)
    return np.concatenate(pts).astype(np.float32), np.concatenate(nrm).astype(np.float32)


def lattice_scene(shape, dx=0.05, jitter=0.2, vel_sigma=0.1, seed=0, origin=(0.0, 0.0, 0.0), open_top=False, wall_dx=None,
                  wall_gap=None, headroom=0):
    """Jittered fluid lattice of ``shape`` = (nx, ny, nz) particles (entries of 1 collapse that axis) inside a box of
    wall particles.  Returns dict(pos, vel, box, box_normals) float32."""
    rng = np.random.default_rng(seed)
    dims = [a for a in range(3) if shape[a] > 1]
    axes = [(np.arange(shape[a]) + 0.5) * dx if a in dims else np.zeros(1) for a in range(3)]
    g = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, 3)
    mask = np.array([1.0 if a in dims else 0.0 for a in range(3)])
    pos = g + rng.uniform(-jitter * dx, jitter * dx, g.shape) * mask + np.asarray(origin)
    vel = rng.normal(0.0, vel_sigma, g.shape) * mask
    lo = np.asarray(origin, dtype=np.float64)
    hi = lo + np.array([shape[a] * dx if a in dims else 0.0 for a in range(3)])
    if headroom:  # the box is taller than the fluid block by `headroom` lattice layers (free surface under an open top)
        hi[1] += headroom * dx
    box, normals = _walls(lo, hi, dx if wall_dx is None else wall_dx, dims, open_top, wall_gap)
    return dict(pos=pos.astype(np.float32), vel=vel.astype(np.float32), box=box, box_normals=normals)


def hydrostatic_scene_2d(n=30, dx=0.005, layers=3, jitter=0.1, seed=3):
    """A resting n x n block of fluid filling a 2-D box whose walls are `layers` rings of wall particles at pitch dx with
    inward normals (the wall sampling of the WBC-SPH data).  Returns dict(pos, vel, acc, box, box_normals)."""
    sc = lattice_scene((n, n, 1), dx=dx, jitter=jitter, vel_sigma=0.0, seed=seed)
    lo, hi = np.zeros(3), np.array([n * dx, n * dx, 0.0])
    rings = [_walls(lo - k * dx, hi + k * dx, dx, [0, 1]) for k in range(layers)]
    sc["box"] = np.concatenate([r[0] for r in rings]).astype(np.float32)
    sc["box_normals"] = np.concatenate([r[1] for r in rings]).astype(np.float32)
    sc["acc"] = np.tile(np.array([[0.0, -9.81, 0.0]], np.float32), (sc["pos"].shape[0], 1))
    return sc


def c4_model_cfg():
    """BASELINE.json config 4: single-scale ASCC+CConv stack (SymNet with strides [1]) on a 3-D box, Liquid3d physics."""
    return dict(name="SymNet", layer_channels=[[[8]], [[32]], [[32]], [[32]], [[3]]], kernel_size=[4, 4, 4],
                sym_kernel_size=[6, 6, 6], coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear",
                window="poly6", window_sym="peak", strides=[1], particle_radii=[0.1], timestep=0.02, grav=-9.81,
                out_scale=[0.0078125] * 3, centralize=True, voxel_size=[0.025] * 3, sym_axis=1, add_merge=True,
                use_acc=False)


def liquid3d_model_cfg():
    """BASELINE.json configs 3 and 5: the full multi-scale net of configs/Liquid3d.yml (three scales, voxel strides 1 / 2 / 4,
    antisymmetric output layer); the shipped checkpoint tests/golden/ckpt_Liquid3d.npz fits it."""
    return dict(name="SymNet", layer_channels=[[[8]], [[16], [8], [4]], [[32], [16], [8]], [[32]], [[3]]],
                kernel_size=[4, 4, 4], sym_kernel_size=[6, 6, 6], coordinate_mapping="ball_to_cube_volume_preserving",
                interpolation="linear", window="poly6", window_sym="peak", window_dens="poly6", strides=[1, 2, 4],
                particle_radii=[0.1, 0.2, 0.4], timestep=0.02, grav=-9.81, out_scale=[0.0078125] * 3, centralize=True,
                voxel_size=[0.025] * 3, sym_axis=1, rest_dens=8.0, circular=False, add_merge=True, use_pre_adv=False,
                use_acc=False, dens_norm=False, dens_feats=False, pres_feats=False)


def slab_scene(n_side, rank, world, dx=0.05, jitter=0.2, vel_sigma=0.1, seed=0, open_top=False):
    """Rank ``rank``'s share of a (world*n_side) x n_side x n_side box split into slabs along x (weak scaling: every
    rank owns n_side^3 fluid particles).  Walls exist on the outer faces only; wall particles are owned by
    coordinate.  Returns (scene dict, slab faces along x)."""
    length = n_side * dx
    sc = lattice_scene((n_side, n_side, n_side), dx=dx, jitter=jitter, vel_sigma=vel_sigma, seed=seed + rank,
                       origin=(rank * length, 0.0, 0.0))
    lo = np.zeros(3)
    hi = np.array([world * length, length, length])
    box, normals = _walls(lo, hi, dx, [0, 1, 2], open_top)
    inf = float("inf")
    faces = [-inf] + [k * length for k in range(1, world)] + [inf]
    own = (box[:, 0] >= faces[rank]) & (box[:, 0] < faces[rank + 1])
    # jitter must not push a fluid particle across a face (ownership is by coordinate)
    sc["pos"][:, 0] = np.clip(sc["pos"][:, 0], rank * length + 1e-4, (rank + 1) * length - 1e-4)
    sc["box"], sc["box_normals"] = box[own], normals[own]
    return sc, faces


def slab_partition(scene, rank, world, n_side, dx=0.05):
    """Strong scaling: rank ``rank``'s share of ONE ``lattice_scene`` with ``n_side`` lattice layers along x, split into ``world``
    slabs whose faces sit on layer boundaries (layers dealt out as evenly as possible: 100 layers on 8 ranks = 13,13,13,13,12,12,12,12;
    the jitter keeps every particle inside its layer, so no fluid particle sits on a face).  Wall particles are owned by
    coordinate.  Returns (scene dict of the rank, slab faces along x)."""
    layers = [n_side // world + (1 if k < n_side % world else 0) for k in range(world)]
    inner = np.cumsum(layers)[:-1] * dx
    inf = float("inf")
    faces = [-inf] + [float(f) for f in inner] + [inf]
    out = {}
    own = (scene["pos"][:, 0] >= faces[rank]) & (scene["pos"][:, 0] < faces[rank + 1])
    out["pos"], out["vel"] = scene["pos"][own], scene["vel"][own]
    own_b = (scene["box"][:, 0] >= faces[rank]) & (scene["box"][:, 0] < faces[rank + 1])
    out["box"], out["box_normals"] = scene["box"][own_b], scene["box_normals"][own_b]
    return out, faces
