"""PBFNet / HRNet / SymNet / CConv with the reference's constructor kwargs and dataflow
(models/base_model.py, models/pbf_model.py, models/hrnet.py, models/sym_net.py, models/cconv.py), on torch CUDA
tensors and the sm_100a kernels.  Inference only.

Two execution modes produce the same numbers (tests assert it):
  * ``fused=False``: every layer is called on its own exactly like the reference does (one neighbour search per
    conv, separate Dense / relu / add): the readable statement of the dataflow;
  * ``fused=True`` (default): what the rollout runs.  Per step: particles are put in cell order for locality, one
    neighbour list per distinct (input set, output set, radius) is built and shared by every conv that needs it,
    relu / window / Dense / bias / residual / add-merge / the antisymmetric centre term ride inside the conv kernel,
    and the two input convs + two input Dense layers collapse into one conv over zero-padded features.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from .convolutions import ContinuousConv, PointSampling, _init
from .losses import compute_density, get_dilated_pos, get_window_func

__all__ = ["BaseModel", "PBFNet", "HRNet", "SymNet", "CConv", "Dense", "align_vector"]


class Dense(torch.nn.Module):
    """tf.keras.layers.Dense(units, activation=None): x @ kernel[Cin,Cout] + bias, built lazily."""

    def __init__(self, units, name=None, activation=None):
        super().__init__()
        assert activation is None
        self.units = int(units)
        self.layer_name = name
        self.kernel = None
        self.bias = None

    def build(self, in_channels, device, generator=None):
        self.kernel = torch.nn.Parameter(_init("glorot_uniform", (int(in_channels), self.units), device,
                                               generator=generator), requires_grad=False)
        self.bias = torch.nn.Parameter(torch.zeros(self.units, device=device), requires_grad=False)

    def forward(self, x, relu_input=False):
        if self.kernel is None:
            self.build(x.shape[-1], x.device)
        if torch.is_grad_enabled() and (self.kernel.requires_grad or self.bias.requires_grad or x.requires_grad):
            return (torch.relu(x) if relu_input else x) @ self.kernel + self.bias  # training path: plain torch ops
        return ops.dense(x, self.kernel, self.bias, relu_input=relu_input)


def align_vector(v0, v1):
    """Rotation taking v0 onto v1 (models/pbf_model.py:12-28); 3x3 float32 tensor on v1's device.  Branch-free (the
    degenerate parallel / antiparallel case is a select on the device): no host sync, CUDA-graph capturable."""
    v0n = v0 / (torch.linalg.norm(v0) + 1e-9)
    v1n = v1 / (torch.linalg.norm(v1) + 1e-9)
    v = torch.linalg.cross(v0n, v1n)
    c = torch.dot(v0n, v1n)
    s = torch.linalg.norm(v)
    eye = torch.eye(3, device=v1.device, dtype=v1.dtype)
    z = torch.zeros((), device=v1.device, dtype=v1.dtype)
    vx = torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])
    general = eye + vx + (vx @ vx) / torch.where(s < 1e-6, torch.ones_like(c), 1 + c)
    return torch.where(s < 1e-6, eye * torch.where(c < 0, -torch.ones_like(c), torch.ones_like(c)), general)


class BaseModel(torch.nn.Module):
    """models/base_model.py:10-29: call = transform -> preprocess -> forward -> postprocess -> inv_transform."""

    def __init__(self, name, **kwargs):
        super().__init__()
        self.model_name = name
        self.cfg = dict(kwargs)

    @property
    def name(self):
        return self.model_name

    def __call__(self, data, training=False, **kwargs):
        return self.call(data, training=training, **kwargs)

    def call(self, data, training=False, **kwargs):
        d = self.transform(data, training=training, **kwargs)
        x = self.preprocess(d, training=training, **kwargs)
        x = self.forward(x, d, training=training, **kwargs)
        x = self.postprocess(x, d, training=training, **kwargs)
        x = self.inv_transform(x, data, training=training, **kwargs)
        return x


class _StepCache:
    """Per-step neighbour lists shared between convs (the reference rebuilds one per conv call)."""

    def __init__(self):
        self.cells = {}
        self.nns = {}
        self.recs = {}

    def search(self, key, inp_pos, out_pos, radius):
        k = (key, float(radius))
        if k not in self.nns:
            ck = (key[0], float(radius))
            if ck not in self.cells:
                self.cells[ck] = ops.CellList(inp_pos, float(radius))
            self.nns[k] = ops.fixed_radius_search(inp_pos, out_pos, radius, ignore_query_point=False,
                                                  return_distances=False, cell_list=self.cells[ck])
        return self.nns[k]

    def records(self, key, nns, kernel_size, inp_pos, out_pos, extent, mapping, interpolation, window, skip_self,
                only_if_shared=False):
        """Pair geometry shared by every conv on the same (neighbour list, filter grid, window).  ``only_if_shared``: a conv whose
        kernel does not need sorted records (k_cconv_direct) and that is the only user of its filter grid evaluates the geometry in
        the kernel instead (no 20 B / pair written and read back); it still takes records another conv has already made."""
        k = (key, tuple(kernel_size), float(extent), mapping, interpolation, window.typ if window else None,
             window.fac if window else 1.0, bool(skip_self))
        if k not in self.recs and only_if_shared:
            return None
        if k not in self.recs:
            self.recs[k] = ops.prepare_pair_records(
                kernel_size, out_pos, extent, None, inp_pos, None, nns.neighbors_index, None, nns.neighbors_row_splits,
                align_corners=True, coordinate_mapping=mapping, interpolation=interpolation,
                window=window.typ if window else None, window_fac=window.fac if window else 1.0, skip_self=skip_self)
        return self.recs[k]


class PBFNet(BaseModel):
    """models/pbf_model.py:31-489."""

    def __init__(self, name="PBFNet", kernel_size=[4, 4, 4], channels=16, strides=[1], particle_radii=[0.05],
                 coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", window=None,
                 window_dens=None, ignore_query_points=False, grav=-9.81, transformation={}, loss=None, timestep=0.01,
                 dens_radius=None, circular=False, dens_feats=False, pres_feats=False, equivar=False, use_vel=True,
                 use_acc=True, use_feats=False, use_box_feats=True, use_pre_adv=False, use_bnds=True, dens_norm=False,
                 rest_dens=3.5, stiffness=20.0, voxel_size=None, centralize=False, out_scale=[0.01, 0.01, 0.01],
                 sample_pad=0, sample_hyst=0.1, part_scale=1.0, fused=True, **kwargs):
        super().__init__(name=name, **kwargs)
        if equivar:
            raise NotImplementedError("equivar=True is disabled in every shipped config and not implemented (SURVEY 8 a16)")
        # Optional input branches (off in every shipped main config): density / pressure input features
        # (models/pbf_model.py:351-367), extra per-particle features, the pre-advection conv (:154-175, 388-399) and the
        # density pyramid for dens_norm (:177-181, 421-435).  They run on the layer-by-layer path (the reference's own
        # sequence of layer calls on the CUDA layers), not on the fused step.
        self.dens_feats, self.pres_feats, self.dens_norm = bool(dens_feats), bool(pres_feats), bool(dens_norm)
        self.use_pre_adv, self.use_feats = bool(use_pre_adv), bool(use_feats)
        if self.dens_feats or self.pres_feats or self.dens_norm or self.use_pre_adv or self.use_feats:
            fused = False
        if voxel_size is None and any(s != 1 for s in strides):
            fused = False  # farthest-point multi-scale sampling (utils/tools/losses.py:274-282): layer-by-layer path
        self.kernel_size = list(kernel_size)
        self.channel = channels
        self.strides = list(strides)
        self.particle_radii = [float(r) for r in particle_radii]
        self.coordinate_mapping = coordinate_mapping
        self.interpolation = interpolation
        self.window = window
        self.window_dens = window_dens
        self.ignore_query_points = ignore_query_points
        self.voxel_size = None if voxel_size is None else [float(v) for v in voxel_size]
        self.centralize = centralize
        self.circular = circular
        self.transformation = dict(transformation or {})
        self.sample_pad = sample_pad
        self.sample_hyst = sample_hyst
        self.out_scale = [float(v) for v in out_scale]
        self.use_vel, self.use_acc, self.use_box_feats, self.use_bnds = use_vel, use_acc, use_box_feats, use_bnds
        self.timestep = float(timestep)
        self.grav = float(grav)
        self.part_scale = float(part_scale)
        self.rest_dens, self.stiffness = rest_dens, stiffness
        self.dens_radius = dens_radius if dens_radius is not None else particle_radii
        self.fused = fused
        self.slab = None  # dmcf_b200.slab.SlabContext for multi-GPU runs (set_slab)
        self.num_fluid_neighbors = None
        self._all_convs = []
        self._box_cache = None
        self._halo = self._pos_in = self._pos_own = None
        self._wcache = {}
        self.fluid_convs = self.get_cconv(name="fluid_obs", filters=channels, window_func=self.window, circular=circular)
        self.fluid_dense = Dense(channels, name="fluid_dense")
        self.obs_convs = self.get_cconv(name="obs_conv", filters=channels, window_func=self.window, circular=circular)
        self.obs_dense = Dense(channels, name="obs_dense")
        if self.use_pre_adv:  # models/pbf_model.py:154-175 (adv_conv1 / adv_dense1 are constructed but never called)
            self.adv_convs = torch.nn.ModuleList([
                self.get_cconv(name="adv_conv0", filters=channels, window_func=self.window, circular=circular),
                self.get_cconv(name="adv_conv1", filters=channels, window_func=self.window, circular=circular)])
            self.adv_dense = torch.nn.ModuleList([Dense(channels, name="adv_dense0"), Dense(channels, name="adv_dense1")])
            self.adv_convs[0]._aliases, self.adv_convs[1]._aliases = ["adv_convs/0"], ["adv_convs/1"]
        if self.dens_norm:  # :177-181
            self.sampling = PointSampling(name="sampling", window_function=get_window_func(self.window_dens), normalize=True)
        self.setup()
        self._convs_ml = torch.nn.ModuleList([c for _, c in self._all_convs])

    def setup(self):
        return

    def set_slab(self, slab):
        """Multi-GPU: this rank owns one spatial slab; convs see [owned | ghost] inputs (dmcf_b200/slab.py)."""
        if slab is not None and slab.world > 1:
            if self.voxel_size is None and any(s != 1 for s in self.strides):
                raise NotImplementedError("slab decomposition needs voxel-grid multi-scale sampling")
            if not self.use_bnds:
                raise NotImplementedError("slab decomposition with use_bnds=False")
            if self.transformation:
                # the slab faces live in the un-transformed frame (migration compares them with inv_transformed positions)
                # while halos / ownership inside the step would see translated / scaled / rotated positions
                raise NotImplementedError("slab decomposition with a non-empty `transformation` (translate / scale / "
                                          "grav_eqvar): ownership and halos would be evaluated in two different frames")
            self.fused = True
        self.slab = slab

    # -- layer factory: models/pbf_model.py:197-224 --------------------------------------------------------
    def get_cconv(self, name, kernel_size=None, activation=None, ignore_query_points=None, window_func=None,
                  normalize=False, **kwargs):
        if kernel_size is None:
            kernel_size = self.kernel_size
        if ignore_query_points is None:
            ignore_query_points = self.ignore_query_points
        conv = ContinuousConv(name=name, kernel_size=kernel_size, activation=activation, align_corners=True,
                              interpolation=self.interpolation, coordinate_mapping=self.coordinate_mapping,
                              normalize=normalize, window_function=get_window_func(window_func),
                              radius_search_ignore_query_points=ignore_query_points, use_dense_layer_for_center=False,
                              **kwargs)
        self._all_convs.append((name, conv))
        return conv

    # -- physics: models/pbf_model.py:234-250 -----------------------------------------------------------------
    def integrate_pos_vel(self, pos1, vel1, acc1=None):
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (pos1, vel1, acc1)):
            # training path (unrolled steps, pipelines/simulator.py:316-421): the same arithmetic in torch ops so that the loss of
            # step t+1 reaches the correction of step t through pos / vel (d pos2 / d pos = I, d pos2 / d vel = dt)
            a = acc1 if acc1 is not None else torch.tensor([0.0, self.grav, 0.0], dtype=pos1.dtype, device=pos1.device)
            vel2 = vel1 + self.timestep * a
            return pos1 + self.timestep * vel2, vel2
        return ops.integrate(pos1, vel1, acc1, (0.0, self.grav, 0.0), self.timestep)

    def compute_new_pos_vel(self, pos1, vel1, pos2, vel2, pos_correction):
        pos = pos2 + pos_correction
        return pos, (pos - pos1) / self.timestep

    # -- transform / inv_transform: models/pbf_model.py:252-301 -----------------------------------------------
    def transform(self, data, training=False, **kwargs):
        pos, vel, acc, feats, box, bfeats = data
        dev = pos.device
        tr = self.transformation
        if "translate" in tr:
            t = self._const("translate", tr["translate"], dev)
            pos, box = pos + t, box + t
        if "scale" in tr:
            s = self._const("scale", tr["scale"], dev)
            pos, box, vel = pos * s, box * s, vel * s
            if acc is not None:
                acc = acc * s
        if "grav_eqvar" in tr:
            g = self._const("grav_eqvar", tr["grav_eqvar"], dev)
            if acc is None or acc.shape[0] == 0:
                raise ValueError("grav_eqvar needs the per-particle acceleration of at least one particle (data[2])")
            self.R = align_vector(g, acc[0])
            pos, vel, acc, box, bfeats = pos @ self.R, vel @ self.R, acc @ self.R, box @ self.R, bfeats @ self.R
        return [pos, vel, acc, feats, box, bfeats]

    def inv_transform(self, prev, data, training=False, **kwargs):
        pos, vel = prev
        dev = pos.device
        tr = self.transformation
        if "grav_eqvar" in tr:
            Rt = self.R.t()
            pos, vel = pos @ Rt, vel @ Rt
        if "scale" in tr:
            s = torch.clamp(self._const("scale", tr["scale"], dev), min=1e-5)
            pos, vel = pos / s, vel / s
        if "translate" in tr:
            pos = pos - self._const("translate", tr["translate"], dev)
        return pos, vel

    def _const(self, name, value, device):
        """Small constant of the config as a device tensor, uploaded once (an upload per step would be a host sync)."""
        key = ("const", name, str(device))
        if key not in self._wcache:
            self._wcache[key] = torch.tensor(value, dtype=torch.float32, device=device)
        return self._wcache[key]

    # -- full step ------------------------------------------------------------------------------------------
    def call(self, data, training=False, **kwargs):
        if training and self.fused:
            raise NotImplementedError("training=True runs on the layer-by-layer path: call set_trainable(True) (or set "
                                      "fused=False) first; the fused step has no backward")
        data = list(data)
        if len(data) != 6:
            raise ValueError("data must be [pos, vel, acc, feats, box, box_normals]")
        perm = None
        n_dev = ops.count_of(data[0]) if self.fused else None  # capacity-sized particle set (sync-free slab rollouts)
        if self.fused and data[0].shape[0] > 0:
            # cell order for locality; undone on the outputs.  Not part of the reference semantics.
            cl = ops.CellList(data[0], self.particle_radii[0])
            perm = cl.sorted_index[: data[0].shape[0]].long()
            if n_dev is not None:  # padding rows stay where they are (their sorted_index entries are not written)
                ar = torch.arange(perm.shape[0], device=perm.device)
                perm = torch.where(ar < n_dev, perm, ar)
            data[0], data[1] = data[0][perm], data[1][perm]
            if data[2] is not None:
                data[2] = data[2][perm]
            if n_dev is not None:
                ops.with_count(data[0], n_dev)
            data[4], data[5] = self._sorted_box(data[4], data[5])
        pos, vel = super().call(data, training=training, **kwargs)
        if perm is not None:
            pos_o, vel_o = torch.empty_like(pos), torch.empty_like(vel)
            pos_o[perm], vel_o[perm] = pos, vel
            pos, vel = pos_o, vel_o
        if n_dev is not None:
            ops.with_count(pos, n_dev)
            ops.with_count(vel, n_dev)
        return pos, vel

    def _sorted_box(self, box, bfeats):
        """The boundary is static over a rollout: put it in cell order once and reuse (a few boundaries are remembered, so
        that two simulators sharing the model -- or alternating scenes -- do not evict each other)."""
        key = (box.data_ptr(), bfeats.data_ptr(), box.shape[0], box._version, bfeats._version)
        if not isinstance(self._box_cache, dict):
            self._box_cache = {}
        hit = self._box_cache.get(key)
        if hit is None:
            if box.shape[0] > 0:
                with ops.no_plan():  # static over the rollout: not one of the step's data-dependent sizes
                    perm = ops.CellList(box, self.particle_radii[0]).sorted_index[: box.shape[0]].long()
                hit = (box[perm].contiguous(), bfeats[perm].contiguous(), box, bfeats)  # keeps the keyed tensors alive
            else:
                hit = (box, bfeats, box, bfeats)
            if len(self._box_cache) >= 4:
                self._box_cache.pop(next(iter(self._box_cache)))
            self._box_cache[key] = hit
        return hit[0], hit[1]

    # -- preprocess: models/pbf_model.py:303-438 ----------------------------------------------------------------
    def preprocess(self, data, training=False, **kwargs):
        _pos, _vel, acc, feats, box, bfeats = data
        n_fluid_dev = ops.count_of(_pos) if self.fused else None
        pos, vel = self.integrate_pos_vel(_pos, _vel, acc)
        if n_fluid_dev is not None:
            ops.with_count(pos, n_fluid_dev)
        filter_extent = [np.float32(r) * np.float32(2) for r in self.particle_radii]
        e_last = float(filter_extent[-1])
        slab = self.slab if (self.slab is not None and self.slab.world > 1) else None
        n_box_dev = None
        if pos.shape[0] > 0 or slab is not None:
            if n_fluid_dev is not None:  # bounding box of the valid rows only
                m = ops.valid_rows_mask(pos)[:, None]
                inf = torch.full((), float("inf"), device=pos.device)
                lo, hi = torch.where(m, pos, inf).amin(dim=0), torch.where(m, pos, -inf).amax(dim=0)
            elif pos.shape[0] > 0:
                lo, hi = pos.amin(dim=0), pos.amax(dim=0)
            else:
                inf = torch.full((3,), float("inf"), device=pos.device)
                lo, hi = inf, -inf
            if slab is not None:  # the cull uses the GLOBAL fluid bounding box (models/pbf_model.py:330-334)
                lo, hi = slab.all_reduce_minmax(lo, hi)
            fltr = ((box >= lo - e_last) & (box <= hi + e_last)).all(dim=1)
            plan = ops.get_plan() if self.fused else None
            if plan is not None and plan.mode == "replay":
                # sync-free cull: stable compaction into a capacity-sized buffer, the count stays on the device
                e, slot = plan.next("rows")
                cap = min(box.shape[0], int(e["n"] * ops.StepPlan.ROW_SLACK) + 256)
                idx, n_box_dev = ops.compact_mask(fltr, cap, plan.hard(slot))
                box, bfeats = ops.with_count(box[idx], n_box_dev), bfeats[idx]
            else:
                box, bfeats = box[fltr], bfeats[fltr]
                if plan is not None:
                    plan.record("rows", n=box.shape[0])
        n_f, n_b = pos.shape[0], box.shape[0]
        fluid_feats = [torch.ones_like(pos[:, :1])]
        if self.use_vel:
            fluid_feats.append(vel)
        if self.use_acc:
            if acc is None:
                raise ValueError("use_acc=True needs the per-particle acceleration (data[2])")
            fluid_feats.append(acc)
        if self.use_feats:
            if feats is None:
                raise ValueError("use_feats=True needs the per-particle features (data[3])")
            fluid_feats.append(feats)
        box_feats = [torch.ones_like(box[:, :1])]
        if self.use_box_feats:
            box_feats.append(bfeats)
        # [fluid | boundary]; with capacity-sized parts the valid rows of both form the valid prefix
        all_pos = ops.concat_rows([pos, box]) if self.fused else torch.cat([pos, box], dim=0)
        self.all_pos = all_pos
        dens0 = None
        if self.dens_feats or self.dens_norm or self.pres_feats:  # models/pbf_model.py:351-367
            win_d = get_window_func(self.window_dens)
            dens0 = compute_density(all_pos, all_pos, float(self.dens_radius[0]), win=win_d)
            if self.dens_feats:
                fluid_feats.append(dens0[:pos.shape[0]].unsqueeze(-1))
                box_feats.append(dens0[pos.shape[0]:].unsqueeze(-1))
            if self.pres_feats:  # utils/tools/losses.py:367-377
                pres = torch.relu(self.stiffness * ((dens0 / self.rest_dens) ** 7 - 1))
                fluid_feats.append(pres[:pos.shape[0]].unsqueeze(-1))
                box_feats.append(pres[pos.shape[0]:].unsqueeze(-1))
        fluid_feats = torch.cat(fluid_feats, dim=-1)
        box_feats = torch.cat(box_feats, dim=-1)
        self.inp_feats, self.inp_bfeats = fluid_feats, box_feats
        self._n_fluid = n_f
        self._step = _StepCache()
        ext0 = float(filter_extent[0])
        ch = self.channel
        if not self.fused:
            ans_conv = self.fluid_convs(fluid_feats * self.part_scale, pos, all_pos, ext0, None)  # :378
            ans_dense = self.fluid_dense(fluid_feats)
            ans_obs = self.obs_convs(box_feats * self.part_scale, box, all_pos, ext0, None)  # :382
            ans_dense_obs = self.obs_dense(box_feats)
            ans_dense = torch.cat([ans_dense, ans_dense_obs], dim=0)
            if self.use_pre_adv:  # :388-399: a third input conv from the positions BEFORE the advection step
                pre_adv_feats = torch.ones_like(_pos[:, :1])
                if self.use_vel:
                    pre_adv_feats = torch.cat([pre_adv_feats, _vel], dim=-1)
                ans_adv = self.adv_convs[0](pre_adv_feats * self.part_scale, _pos, all_pos, ext0, None)
                ans_dens_adv = torch.cat([self.adv_dense[0](pre_adv_feats), ans_dense_obs], dim=0)
                feats_out = torch.cat([ans_conv, ans_obs, ans_adv, ans_dense, ans_dens_adv], dim=-1)
            else:
                feats_out = torch.cat([ans_conv, ans_obs, ans_dense], dim=-1)  # :411
        else:
            cf, cb = fluid_feats.shape[1], box_feats.shape[1]
            self._ensure_built_inputs(cf, cb, pos.device)
            if ops.count_of(all_pos) is None:
                x = torch.zeros((n_f + n_b, cf + cb), dtype=torch.float32, device=pos.device)
                x[:n_f, :cf] = fluid_feats
                x[n_f:, cf:] = box_feats
            else:  # capacity-sized parts: pad each block to the full width, then join their valid rows
                xf = torch.cat([fluid_feats, torch.zeros((n_f, cb), dtype=torch.float32, device=pos.device)], dim=1)
                xb = torch.cat([torch.zeros((n_b, cf), dtype=torch.float32, device=pos.device), box_feats], dim=1)
                x = ops.concat_rows([ops.with_count(xf, n_fluid_dev) if n_fluid_dev is not None else xf,
                                     ops.with_count(xb, n_box_dev) if n_box_dev is not None else xb])
            w, b = self._input_weights(cf, cb)
            all_in = all_pos
            self._halo = None
            self._pos_own = [all_pos]
            if slab is not None:  # ghosts of the two neighbouring slabs: positions once per step, features per layer
                # halo width = the largest radius any conv applies to these points (coarse scales read scale 0 with it)
                self._halo = [slab.make_halo(all_pos, max(self.particle_radii))]
                all_in = ops.concat_rows([all_pos, self._halo[0].ghost_pos])
                # the owned rows as a VIEW of [owned | ghost]: same values, and a search whose queries are a prefix of its points
                # takes the cell-centric kernel (dmcf_frs_*: queries == grid->points)
                cnt = ops.count_of(all_pos)
                all_pos = all_in[: all_pos.shape[0]]
                if cnt is not None:
                    ops.with_count(all_pos, cnt)
                self.all_pos = all_pos
                self._pos_own = [all_pos]
                self._pos_in = [all_in]
                x = self._with_ghosts(0, x)
            self._pos_in = [all_in]
            nns = self._step.search((0, 0), all_in, all_pos, 0.5 * ext0)
            win = self.fluid_convs.window_function
            recs = self._step.records((0, 0), nns, self.kernel_size, all_in, all_pos, ext0, self.coordinate_mapping,
                                      self.interpolation, win, self.ignore_query_points)
            feats_out, own_rows = self._feature_buffer(0, all_pos.shape[0], w.shape[1], pos.device)
            ops.continuous_conv(
                w, all_pos, ext0, None, all_in, x, None, nns.neighbors_index, None, nns.neighbors_row_splits,
                align_corners=True, coordinate_mapping=self.coordinate_mapping, normalize=False,
                interpolation=self.interpolation, window=win.typ if win else None, window_fac=win.fac if win else 1.0,
                feat_scale=self.part_scale, skip_self=self.ignore_query_points, bias=b, dense_inp=x,
                dense_cin=cf + cb, kernel_size=self.kernel_size, pair_records=recs, out=own_rows,
                block_diagonal=(cf, ch, ch))  # rows are [fluid | 0] or [0 | box], the filter is built block diagonal
        src = all_pos if self.use_bnds else pos
        if slab is not None and self.fused:
            dilated_pos, idx = self._slab_dilated_pos(slab, all_pos, all_in), [None] * len(self.strides)
        else:
            dilated_pos, _, idx = get_dilated_pos(src, self.strides, voxel_size=self.voxel_size,
                                                  centralize=self.centralize, pad=self.sample_pad, hyst=self.sample_hyst)
        self.dilated_pos = dilated_pos
        dens = None
        if self.dens_norm:  # :421-435: density pyramid, resampled scale to scale (the radius is passed as the extent)
            dens = [(dens0 if self.use_bnds else dens0[:n_f]).unsqueeze(-1)]
            for sc in range(1, len(self.dens_radius)):
                d = self.sampling(dens[-1], dilated_pos[sc - 1], dilated_pos[sc], float(self.dens_radius[sc]), None)
                dens.append(torch.clamp(d, min=1e-2))
        return [dilated_pos, feats_out, idx, dens]

    def _with_ghosts(self, scale, x):
        """Layer input of one scale under slab decomposition: [owned rows | ghost rows refreshed from the neighbours].  A
        buffer allocated by ``_feature_buffer`` has room for the ghost rows behind its owned rows and is filled in place;
        anything else is copied into a new [owned | ghost] array."""
        own = self._pos_own[scale]
        cnt = ops.count_of(own)
        if x.shape[0] == self._pos_in[scale].shape[0] and x.shape[0] > own.shape[0]:
            return self._halo[scale].fill_ghosts(x, cnt if cnt is not None else int(own.shape[0]))
        if cnt is not None and ops.count_of(x) is None:
            ops.with_count(x, cnt)  # feature rows follow the owned points of their scale
        return self._halo[scale].with_ghosts(x)

    def _feature_buffer(self, scale, n_rows, channels, device):
        """Output buffer of a conv on the points of ``scale``: under slab decomposition with room for the ghost rows of the
        next layer's input behind the ``n_rows`` owned rows (returns (whole buffer, view of the owned rows))."""
        rows = n_rows
        if self.slab is not None and self.slab.world > 1 and self._halo is not None and scale < len(self._pos_in) \
                and self._pos_in[scale] is not None:
            rows = max(n_rows, int(self._pos_in[scale].shape[0]))
        buf = torch.empty((rows, channels), dtype=torch.float32, device=device)
        return buf, buf[:n_rows]

    def _slab_dilated_pos(self, slab, all_own, all_in):
        """Multi-scale lattices under slab decomposition: every rank builds the lattice from its owned + ghost particles
        (all particles within one coarse voxel of the slab are among the ghosts), keeps the lattice points whose
        coordinate falls in its slab (ownership by coordinate makes the per-rank sets a partition of the global
        lattice) and receives the neighbours' lattice points near the faces as ghosts of that scale."""
        from .losses import grid_pos
        out = []
        center = None
        for si, stride in enumerate(self.strides):
            if stride == 1:
                out.append(all_own)
                continue
            if self.centralize and center is None:  # global mean, float64 accumulate, rounded once (losses.point_mean)
                m = ops.valid_rows_mask(all_own)
                p64 = all_own.to(torch.float64)
                if m is not None:
                    p64 = torch.where(m[:, None], p64, torch.zeros((), dtype=torch.float64, device=all_own.device))
                    n64 = ops.count_of(all_own).to(torch.float64)
                else:
                    n64 = torch.full((1,), float(all_own.shape[0]), dtype=torch.float64, device=all_own.device)
                acc = slab.all_reduce_sum(torch.cat([p64.sum(dim=0), n64]))
                center = (acc[:3] / acc[3]).to(torch.float32)
            vs = torch.as_tensor(self.voxel_size, dtype=torch.float32) * float(stride)
            lat = grid_pos(all_in, vs, self.centralize, self.sample_pad, self.sample_hyst, center)
            (own,) = slab._select(slab.owned_mask(lat), lat)
            own = own.contiguous() if ops.count_of(own) is None else own
            plan = slab.make_halo(own, max(self.particle_radii))
            while len(self._halo) <= si:
                self._halo.append(None)
                self._pos_in.append(None)
                self._pos_own.append(None)
            self._halo[si] = plan
            self._pos_in[si] = ops.concat_rows([own, plan.ghost_pos])
            cnt = ops.count_of(own)
            own = self._pos_in[si][: own.shape[0]]  # view of [owned | ghost] (prefix searches, see preprocess)
            if cnt is not None:
                ops.with_count(own, cnt)
            self._pos_own[si] = own
            out.append(own)
        return out

    def _ensure_built_inputs(self, cf, cb, device):
        if self.fluid_convs.kernel is None:
            self.fluid_convs.build(cf, device)
        if self.obs_convs.kernel is None:
            self.obs_convs.build(cb, device)
        if self.fluid_dense.kernel is None:
            self.fluid_dense.build(cf, device)
        if self.obs_dense.kernel is None:
            self.obs_dense.build(cb, device)

    def _input_weights(self, cf, cb):
        """One filter for [fluid conv | obstacle conv | fluid/obstacle Dense] over zero-padded [fluid|box] features.
        Dense biases ride on the constant-one feature channels (0 and cf)."""
        layers = (self.fluid_convs.kernel, self.obs_convs.kernel, self.fluid_dense.kernel, self.obs_dense.kernel,
                  self.fluid_convs.bias, self.obs_convs.bias, self.fluid_dense.bias, self.obs_dense.bias)
        key = tuple((t.data_ptr(), t._version) for t in layers)
        hit = self._wcache.get("input")
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        ch = self.channel
        kf, ko = self.fluid_convs.effective_kernel(), self.obs_convs.effective_kernel()
        cells = kf.shape[0] * kf.shape[1] * kf.shape[2]
        dev = kf.device
        w = torch.zeros((cells, cf + cb, 3 * ch), device=dev)
        w[:, :cf, :ch] = kf.reshape(cells, cf, ch)
        w[:, cf:, ch:2 * ch] = ko.reshape(cells, cb, ch)
        wd = torch.zeros((cf + cb, 3 * ch), device=dev)
        wd[:cf, 2 * ch:] = self.fluid_dense.kernel
        wd[cf:, 2 * ch:] = self.obs_dense.kernel
        wd[0, 2 * ch:] += self.fluid_dense.bias
        wd[cf, 2 * ch:] += self.obs_dense.bias
        b = torch.zeros(3 * ch, device=dev)
        b[:ch] = self.fluid_convs.bias
        b[ch:2 * ch] = self.obs_convs.bias
        w_ext = torch.cat([w.reshape(-1, 3 * ch), wd], dim=0).contiguous()
        self._wcache["input"] = (key, w_ext, b)
        return w_ext, b

    def _block_weights(self, conv, dense):
        """Flattened conv filter with the Dense kernel appended, and the summed bias."""
        ts = [conv.kernel] + ([conv.bias] if conv.bias is not None else []) + ([dense.kernel, dense.bias] if dense else [])
        key = tuple((t.data_ptr(), t._version) for t in ts)
        ck = (id(conv), id(dense))
        hit = self._wcache.get(ck)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        k = conv.effective_kernel()
        w = k.reshape(-1, k.shape[-1])
        b = conv.bias
        if dense is not None:
            w = torch.cat([w, dense.kernel], dim=0)
            b = dense.bias if b is None else b + dense.bias
        w = w.contiguous()
        self._wcache[ck] = (key, w, b)
        return w, b

    def conv_block(self, conv, dense, x, inp_pos, out_pos, extent, key, *, relu=True, scale=1.0, same_set=False,
                   residual=None, out=None, accumulate=False, ascc=False, inp_scale=0):
        """relu -> conv (+ Dense on the same relu'd features) (+ residual), the unit models/hrnet.py:81-99 and
        models/cconv.py:60-67 repeat; fused into one kernel launch."""
        if conv.kernel is None:
            conv.build(x.shape[1], x.device)
        if dense is not None and dense.kernel is None:
            dense.build(x.shape[1], x.device)
        w, b = self._block_weights(conv, dense)
        if self.slab is not None and self.slab.world > 1:
            # owned rows out, [owned | ghost] rows in: refresh the ghost rows of this layer's input from the neighbours
            x = self._with_ghosts(inp_scale, x)
            inp_pos = self._pos_in[inp_scale]
        nns = self._step.search(key, inp_pos, out_pos, 0.5 * float(extent))
        win = conv.window_function
        skip = bool(conv.radius_search_ignore_query_points and same_set)
        direct = ops.conv_kernel_name(conv.kernel_size, x.shape[1], w.shape[1], self.interpolation,
                                      x.shape[1] if dense is not None else 0) == "k_cconv_direct"
        recs = self._step.records(key, nns, conv.kernel_size, inp_pos, out_pos, float(extent), self.coordinate_mapping,
                                  self.interpolation, win, skip, only_if_shared=direct)
        return ops.continuous_conv(
            w, out_pos, float(extent), None, inp_pos, x, None, nns.neighbors_index, None, nns.neighbors_row_splits,
            align_corners=True, coordinate_mapping=self.coordinate_mapping, normalize=False,
            interpolation=self.interpolation, window=win.typ if win else None, window_fac=win.fac if win else 1.0,
            relu_input=relu, feat_scale=scale, ascc=ascc,
            skip_self=skip, bias=b,
            dense_inp=x if dense is not None else None, dense_cin=x.shape[1] if dense is not None else 0,
            residual=residual, out=out, accumulate=accumulate, kernel_size=conv.kernel_size, pair_records=recs,
            antisymmetric_filter=bool(getattr(conv, "symmetric", False)) and not getattr(conv, "circular", False))

    # -- postprocess: models/pbf_model.py:440-489 -----------------------------------------------------------------
    def postprocess(self, prev, data, training=False, **kwargs):
        pos, vel, acc = data[:3]
        pcnt = pos.shape[0]
        out = prev
        self.net_out = out
        pos2, vel2 = self.integrate_pos_vel(pos, vel, acc)
        scale = self.out_scale
        if torch.is_grad_enabled() and out.requires_grad:  # training path: the same arithmetic in torch ops (:466-487)
            o = out.repeat(1, 3) if out.shape[-1] == 1 else (torch.cat([out, out[:, :1]], dim=-1) if out.shape[-1] == 2 else out)
            pos_new = pos2 + torch.tensor(scale, dtype=torch.float32, device=out.device) * o[:pcnt]
            return [pos_new, (pos_new - pos) / self.timestep]
        pos_new, vel_new = ops.correct(pos, pos2, out, scale, self.timestep)
        return [pos_new, vel_new]

    # -- training hooks: models/pbf_model.py:491-517 --------------------------------------------------------------
    def set_trainable(self, flag=True):
        """Training runs on the layer-by-layer path through ``dmcf_b200.autograd`` (layers must be built: load or init
        weights first)."""
        if flag:
            self.fused = False
        for prm in self.parameters():
            prm.requires_grad_(flag)
        self._wcache = {}
        return self

    def loss(self, results, data, loss_fn=None):
        """``results`` = [pos, vel] of a model call, ``data`` = [inputs, target, target_prev, pre_steps] (:494-509)."""
        from .losses import get_loss
        fns = loss_fn if loss_fn is not None else (getattr(self, "loss_fn", None) or {"mse": get_loss("mse")})
        return {n: l(data[1], results[0], num_fluid_neighbors=self.fluid_neighbor_counts(), input=data[0],
                     target_prev=data[2], pre_steps=data[3], pos_correction=None) for n, l in fns.items()}

    def get_optimizer(self, cfg):
        """Adam(eps=1e-6) with the piecewise-constant learning-rate schedule of the YAML (:511-517)."""
        bounds, values = list(cfg["lr_boundaries"]), list(cfg["lr_values"])
        opt = torch.optim.Adam([p for p in self.parameters() if p.requires_grad], lr=values[0], eps=1e-6)
        import bisect
        sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda it: values[bisect.bisect_left(bounds, it)] / values[0])
        return opt, sched

    @property
    def pos_correction(self):
        out = self.net_out
        if out.shape[-1] == 1:
            out = out.repeat(1, 3)
        elif out.shape[-1] == 2:
            out = torch.cat([out, out[:, :1]], dim=-1)
        return torch.tensor(self.out_scale, device=out.device) * out[: self._n_fluid]

    def fluid_neighbor_counts(self):
        """num_fluid_neighbors of models/pbf_model.py:450-453 (training-loss input), computed on demand."""
        n_f = self._n_fluid
        counts, _ = ops.neighbor_counts(self.all_pos[:n_f], self.all_pos, self.particle_radii[0],
                                        ignore_query_point=self.ignore_query_points)
        return counts[:n_f].to(torch.float32)

    # -- weights --------------------------------------------------------------------------------------------
    def named_layers(self):
        """{checkpoint-style name: layer}."""
        out = {"fluid_convs": self.fluid_convs, "obs_convs": self.obs_convs, "fluid_dense": self.fluid_dense,
               "obs_dense": self.obs_dense}
        for n, (_, conv) in enumerate(self._all_convs):
            if n >= 2:
                out["_all_convs/%d" % n] = conv
        if self.use_pre_adv:
            out["adv_dense/0"], out["adv_dense/1"] = self.adv_dense[0], self.adv_dense[1]
        return out

    def load_weights(self, weights, device="cuda"):
        """Assigns ``{name/kernel|bias: array}`` (see ``dmcf_b200.checkpoint.model_weights``); names follow the TF
        object graph of the shipped checkpoints, '/1' of '_all_convs/<n>/1' being optional."""
        weights = {k.replace("/1/", "/") if k.startswith("_all_convs/") else k: v for k, v in weights.items()}
        missing = []
        for name, layer in self.named_layers().items():
            aliases = [name] + list(getattr(layer, "_aliases", []))
            found = next((a for a in aliases if a + "/kernel" in weights), None)
            if found is None:
                missing.append(name)
                continue
            k = torch.as_tensor(np.asarray(weights[found + "/kernel"]), dtype=torch.float32).to(device)
            layer.kernel = torch.nn.Parameter(k.contiguous(), requires_grad=False)
            has_bias = found + "/bias" in weights
            if isinstance(layer, ContinuousConv):
                layer.in_channels = k.shape[-2]
                if tuple(k.shape) != tuple(layer.kernel_shape(k.shape[-2])):
                    raise ValueError(f"{name}: checkpoint kernel {tuple(k.shape)} does not match the layer "
                                     f"{layer.kernel_shape(k.shape[-2])}")
                layer._eff_cache = None
                if layer.use_bias != has_bias:
                    raise ValueError(f"{name}: bias presence differs between checkpoint and layer")
            if has_bias:
                layer.bias = torch.nn.Parameter(
                    torch.as_tensor(np.asarray(weights[found + "/bias"]), dtype=torch.float32).to(device),
                    requires_grad=False)
        self._wcache = {}
        return missing

    def init_weights(self, seed=0, device="cuda", scale=None):
        """Random weights of the architecture's shapes (no checkpoint): Keras defaults, or uniform(-scale, scale)."""
        gen = torch.Generator().manual_seed(seed)
        for name, layer, cin in self.layer_shapes():
            if isinstance(layer, ContinuousConv):
                layer.build(cin, device, generator=gen)
                if scale is not None:
                    layer.kernel.data = ((torch.rand(layer.kernel.shape, generator=gen) * 2 - 1) * scale).to(device)
                if layer.bias is not None and scale is not None:
                    layer.bias.data = ((torch.rand(layer.bias.shape, generator=gen) * 2 - 1) * scale).to(device)
            else:
                layer.build(cin, device, generator=gen)
                if scale is not None:
                    layer.bias.data = ((torch.rand(layer.bias.shape, generator=gen) * 2 - 1) * scale).to(device)
        self._wcache = {}

    def input_channels(self):
        cf = 1 + (3 if self.use_vel else 0) + (3 if self.use_acc else 0)
        cb = 1 + (3 if self.use_box_feats else 0)
        return cf, cb

    def layer_shapes(self):
        """[(name, layer, in_channels)] for every layer that gets weights (what lazy building would produce)."""
        cf, cb = self.input_channels()
        return [("fluid_convs", self.fluid_convs, cf), ("obs_convs", self.obs_convs, cb),
                ("fluid_dense", self.fluid_dense, cf), ("obs_dense", self.obs_dense, cb)]

    def state_arrays(self):
        """{name/kernel|bias: ndarray} of every built layer (the oracle consumes this)."""
        out = {}
        for name, layer in self.named_layers().items():
            if layer.kernel is not None:
                out[name + "/kernel"] = layer.kernel.detach().cpu().numpy()
                if layer.bias is not None:
                    out[name + "/bias"] = layer.bias.detach().cpu().numpy()
        return out


class HRNet(PBFNet):
    """models/hrnet.py:12-133 (multi-scale conv stack)."""

    def __init__(self, name="HRNet", layer_channels=[[16], [32], [32], [3]], window=None, window_dens=None,
                 circular=False, add_merge=False, out_activation=None, **kwargs):
        self.layer_channels = layer_channels
        self.add_merge = add_merge
        if out_activation == "tanh":
            self.out_activation = torch.tanh
        elif out_activation is None:
            self.out_activation = None
        else:
            raise NotImplementedError()
        torch.nn.Module.__init__(self)
        super().__init__(name=name, channels=layer_channels[0][0][0], window=window, window_dens=window_dens,
                         circular=circular, **kwargs)

    def setup(self):  # models/hrnet.py:39-67
        self.convs, self.denses = [], []
        lc = self.layer_channels
        for i in range(1, len(lc)):
            self.denses.append([])
            self.convs.append([])
            for j in range(len(lc[i])):
                self.convs[-1].append([])
                self.denses[-1].append([])
                for k in range(len(lc[i][j])):
                    ch = lc[i][j][k]
                    self.convs[-1][-1].append([])
                    self.denses[-1][-1].append([])
                    for l in range(len(lc[i - 1]) if k == 0 else 1):
                        conv = self.get_cconv(name="conv{0}{1}{2}_{3}".format(i, j, k, l), filters=ch,
                                              window_func=self.window,
                                              ignore_query_points=self.ignore_query_points and (j == l or k > 0),
                                              circular=self.circular)
                        conv._aliases = ["convs/%d/%d/%d/%d" % (i - 1, j, k, l)]
                        self.convs[-1][-1][-1].append(conv)
                        self.denses[-1][-1][-1].append(Dense(ch, name="dense{0}{1}{2}_{3}".format(i, j, k, l)))
        self._dense_ml = torch.nn.ModuleList([d for a in self.denses for b in a for c in b for d in c])

    def named_layers(self):
        out = super().named_layers()
        for i, a in enumerate(self.denses):
            for j, b in enumerate(a):
                for k, c in enumerate(b):
                    for l, d in enumerate(c):
                        if self._dense_used(i, j, k, l):
                            out["denses/%d/%d/%d/%d" % (i, j, k, l)] = d
        return out

    def _dense_used(self, i, j, k, l):
        # voxel mode: only the diagonal Dense layers ever run (models/hrnet.py:94-99); with farthest-point sampling the
        # cross-scale ones act on gathered / scattered rows (:100-113)
        return k > 0 or j == l or self._fps_scales()

    def _set_key(self, inp_scale, out_scale):
        """Step-cache key of the (input set, output set) pair of a conv: scale 0 is [fluid | boundary] with use_bnds and the
        fluid rows alone without (preprocess stores the all->all list under (0, 0): the two must not collide)."""
        if self.use_bnds:
            return (inp_scale, out_scale)
        return (("f", inp_scale), ("f", out_scale))

    def _fps_scales(self):
        return self.voxel_size is None and any(s != 1 for s in self.strides)

    def _scale_channels(self):
        """Channel count of ans_convs[layer][scale] for every layer (index 0 = the preprocess output)."""
        lc = self.layer_channels
        chans = [[3 * self.channel]]
        for i in range(1, len(lc)):
            cur = []
            for j in range(len(lc[i])):
                c = lc[i][j][0] if self.add_merge else lc[i][j][0] * len(lc[i - 1])
                if len(lc[i][j]) > 1:
                    c = lc[i][j][-1]
                cur.append(c)
            chans.append(cur)
        return chans

    def layer_shapes(self):
        out = super().layer_shapes()
        lc = self.layer_channels
        chans = self._scale_channels()
        for i in range(1, len(lc)):
            for j in range(len(lc[i])):
                for k in range(len(lc[i][j])):
                    for l in range(len(lc[i - 1]) if k == 0 else 1):
                        if k == 0:
                            cin = chans[i - 1][l]
                        elif k == 1:
                            cin = lc[i][j][0] if self.add_merge else lc[i][j][0] * len(lc[i - 1])
                        else:
                            cin = lc[i][j][k - 1]
                        out.append(("conv", self.convs[i - 1][j][k][l], cin))
                        if self._dense_used(i - 1, j, k, l):
                            out.append(("dense", self.denses[i - 1][j][k][l], cin))
        return out

    def forward(self, prev, data, training=False, **kwargs):  # models/hrnet.py:69-133
        pos, feats, idx, dens = prev
        if not self.use_bnds:
            feats = feats[: pos[0].shape[0]]
        filter_extent = [float(np.float32(r) * np.float32(2)) for r in self.particle_radii]
        ans_convs = [[feats]]
        for layer in range(len(self.convs)):
            ans = []
            for scale in range(len(self.convs[layer])):
                importance = self.part_scale if scale == 0 else 1.0
                n_inp = len(ans_convs[-1])
                ext = filter_extent[scale]
                if self.fused:
                    cw = self.convs[layer][scale][0][0].filters
                    n_rows = pos[scale].shape[0]
                    buf, own_rows = self._feature_buffer(scale, n_rows, cw if self.add_merge else cw * n_inp, feats.device)
                    for inp_scale in range(n_inp):
                        x = ans_convs[-1][inp_scale]
                        ext = filter_extent[max(inp_scale, scale)]
                        same = scale == inp_scale
                        res = ans_convs[-1][scale] if same and cw == ans_convs[-1][scale].shape[-1] else None
                        if res is not None:
                            res = res[:n_rows]
                        o = own_rows if self.add_merge else own_rows[:, inp_scale * cw:(inp_scale + 1) * cw]
                        self.conv_block(self.convs[layer][scale][0][inp_scale],
                                        self.denses[layer][scale][0][inp_scale] if same else None, x, pos[inp_scale],
                                        pos[scale], ext, self._set_key(inp_scale, scale), relu=True, scale=importance,
                                        same_set=same, residual=res, out=o,
                                        accumulate=self.add_merge and inp_scale > 0, inp_scale=inp_scale)
                    ans.append(buf)
                else:
                    inp = []
                    for inp_scale in range(n_inp):
                        f = torch.relu(ans_convs[-1][inp_scale])
                        ext = filter_extent[max(inp_scale, scale)]
                        if self.dens_norm and inp_scale < len(dens):  # models/hrnet.py:87-89
                            f = torch.cat([f, f / dens[inp_scale] ** 2], dim=-1)
                        a = self.convs[layer][scale][0][inp_scale](f * importance, pos[inp_scale], pos[scale], ext, None)
                        if scale == inp_scale:
                            a = a + self.denses[layer][scale][0][inp_scale](f)
                            if a.shape[-1] == ans_convs[-1][scale].shape[-1]:
                                a = a + ans_convs[-1][scale]
                        elif self.voxel_size is None:  # models/hrnet.py:100-113: nested farthest-point subsets
                            dense = self.denses[layer][scale][0][inp_scale]
                            if scale > inp_scale:  # the coarse points' own rows of the finer features
                                for i in range(inp_scale, scale):
                                    f = f[idx[i + 1][0].long()]
                                a = a + dense(f)
                            else:  # coarse features added onto the fine points they were sampled from
                                ind = idx[scale + 1][0].long()
                                for i in range(scale + 1, inp_scale):
                                    ind = ind[idx[i + 1][0].long()]
                                a = a.index_add(0, ind, dense(f))
                        inp.append(a)
                    if self.add_merge:
                        s = inp[0]
                        for a in inp[1:]:
                            s = s + a
                        ans.append(s)
                    else:
                        ans.append(torch.cat(inp, dim=-1))
                for i in range(1, len(self.convs[layer][scale])):  # extra k>0 convs on the same scale (:120-129)
                    x = ans[-1]
                    res = None
                    if len(ans_convs[-1]) > scale and self.convs[layer][scale][i][0].filters == ans_convs[-1][scale].shape[-1]:
                        res = ans_convs[-1][scale]
                    if self.fused:
                        ans[-1] = self.conv_block(self.convs[layer][scale][i][0], self.denses[layer][scale][i][0], x,
                                                  pos[scale], pos[scale], ext, self._set_key(scale, scale), relu=False,
                                                  scale=importance, same_set=True, residual=res, inp_scale=scale)
                    else:
                        a = self.convs[layer][scale][i][0](x * importance, pos[scale], pos[scale], ext, None)
                        a = a + self.denses[layer][scale][i][0](x)
                        ans[-1] = a + res if res is not None else a
            ans_convs.append(ans)
        out = ans_convs[-1][0]
        return self.out_activation(out) if self.out_activation is not None else out


class SymNet(HRNet):
    """models/sym_net.py:12-69: HRNet followed by antisymmetric (momentum conserving) ContinuousConv layer(s)."""

    def __init__(self, name="SymNet", layer_channels=[[[16]], [[32]], [[32]], [[3]]], sym_kernel_size=[6, 6, 6],
                 sym_axis=2, window_sym=None, out_activation=None, **kwargs):
        self.sym_kernel_size = list(sym_kernel_size)
        self.sym_axis = sym_axis
        self.window_sym = window_sym
        self.sym_channels = layer_channels[-1][-1]
        if out_activation == "tanh":
            self.act = torch.tanh
        elif out_activation is None:
            self.act = None
        else:
            raise NotImplementedError()
        super().__init__(name=name, layer_channels=layer_channels[:-1], out_activation=None, **kwargs)

    def setup(self):  # models/sym_net.py:39-53
        super().setup()
        self.sym_convs = []
        for i, ch in enumerate(self.sym_channels):
            conv = self.get_cconv(name="sym_conv{0}".format(i), filters=ch, use_bias=False, symmetric=True,
                                  kernel_size=self.sym_kernel_size, ignore_query_points=True,
                                  window_func=self.window_sym, sym_axis=self.sym_axis, circular=self.circular)
            conv._aliases = ["sym_convs/%d" % i]
            self.sym_convs.append(conv)

    def layer_shapes(self):
        out = super().layer_shapes()
        cin = self._scale_channels()[-1][0]
        for conv in self.sym_convs:
            out.append(("sym", conv, cin))
            cin = conv.filters
        return out

    def forward(self, prev, data, training=False, **kwargs):  # models/sym_net.py:55-69
        pos, feats, idx, dens = prev
        ans = super().forward(prev, data, training, **kwargs)
        if not self.use_bnds:
            ans = torch.cat([ans, feats[pos[0].shape[0]:]], dim=0)
        ext = float(np.float32(self.particle_radii[0]) * np.float32(2))
        for conv in self.sym_convs:
            if self.fused:
                ans = self.conv_block(conv, None, ans, self.all_pos, self.all_pos, ext, (0, 0) if self.use_bnds else ("all", "all"),
                                      relu=True, scale=self.part_scale, same_set=True, ascc=True)
            else:
                ans = conv(torch.relu(ans) * self.part_scale, self.all_pos, self.all_pos, ext, None)
        return self.act(ans) if self.act is not None else ans


class CConv(PBFNet):
    """models/cconv.py:12-69 (single-scale baseline of Ummenhofer et al.)."""

    def __init__(self, name="CConv", layer_channels=[32, 64, 64, 3], window=None, out_activation=None, **kwargs):
        self.layer_channels = layer_channels
        if out_activation == "tanh":
            self.out_activation = torch.tanh
        elif out_activation is None:
            self.out_activation = None
        else:
            raise NotImplementedError()
        torch.nn.Module.__init__(self)
        super().__init__(name=name, channels=layer_channels[0], window=window, **kwargs)

    def setup(self):  # models/cconv.py:33-48
        self.convs, self.denses = [], []
        for i in range(1, len(self.layer_channels)):
            ch = self.layer_channels[i]
            conv = self.get_cconv(name="conv{0}".format(i), filters=ch, window_func=self.window,
                                  ignore_query_points=self.ignore_query_points, circular=self.circular)
            conv._aliases = ["convs/%d" % (i - 1)]
            self.convs.append(conv)
            self.denses.append(Dense(ch, name="dense{0}".format(i)))
        self._dense_ml = torch.nn.ModuleList(self.denses)

    def named_layers(self):
        out = super().named_layers()
        for i, d in enumerate(self.denses):
            out["denses/%d" % i] = d
        return out

    def layer_shapes(self):
        out = super().layer_shapes()
        cin = 3 * self.channel
        for conv, dense in zip(self.convs, self.denses):
            out.append(("conv", conv, cin))
            out.append(("dense", dense, cin))
            cin = conv.filters
        return out

    def forward(self, prev, data, training=False, **kwargs):  # models/cconv.py:50-69
        pos, feats = prev[:2]
        pos = pos[0]
        feats = feats[: pos.shape[0]]
        ext = float(np.float32(self.particle_radii[0]) * np.float32(2))
        ans = feats
        key = (0, 0) if self.use_bnds else ("fluid", "fluid")
        for conv, dense in zip(self.convs, self.denses):
            if self.fused:
                res = ans if conv.filters == ans.shape[-1] else None
                ans = self.conv_block(conv, dense, ans, pos, pos, ext, key, relu=True, scale=1.0, same_set=True,
                                      residual=res)
            else:
                f = torch.relu(ans)
                a = conv(f, pos, pos, ext, None) + dense(f)
                ans = a + ans if a.shape[-1] == ans.shape[-1] else a
        return self.out_activation(ans) if self.out_activation is not None else ans
