"""Run-length statistics of the pair records of a step (Liquid3d by default): how many distinct base cells a 32-pair chunk of a
neighbour row holds (the chunks are what k_cconv_prepare sorts and the register-patch walk merges over).
python scripts/run_stats.py [liquid3d|c4]"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from dmcf_b200 import ops, config, scenes
from dmcf_b200.simulator import Simulator
import test_models_gpu as T
which = sys.argv[1] if len(sys.argv) > 1 else 'liquid3d'
dev = torch.device('cuda')
if which == 'c4':
    cfg, scene, wname = scenes.c4_model_cfg(), scenes.lattice_scene((40, 40, 40), dx=0.05, seed=0), None
else:
    cfg, scene, wname = T.liquid3d_cfg(), scenes.lattice_scene((46, 46, 46), dx=0.05, seed=2, open_top=True), 'ckpt_Liquid3d.npz'
model = config.build_model(cfg)
if wname is None:
    model.init_weights(seed=0, device=dev, scale=0.1)
else:
    model.load_weights(T.load_npz_weights(wname), device=dev)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
sample = [t(scene['pos']), t(scene['vel']), None, None, t(scene['box']), t(scene['box_normals'])]
orig = ops.prepare_pair_records


def wrapped(kernel_size, out_positions, extents, offset, inp_positions, inp_importance, neighbors_index, neighbors_importance,
            neighbors_row_splits, **kw):
    rec = orig(kernel_size, out_positions, extents, offset, inp_positions, inp_importance, neighbors_index, neighbors_importance,
               neighbors_row_splits, **kw)
    rs = neighbors_row_splits
    n_out = rs.shape[0] - 1
    P = rec.shape[1]
    row = rec[0].view(torch.int32)
    i0 = rec[1].view(torch.int32)
    kz, ky, kx = kernel_size
    bx = torch.clamp(i0 & 0xff, max=max(kx - 1, 1) - 1); by = torch.clamp((i0 >> 8) & 0xff, max=max(ky - 1, 1) - 1)
    bz = torch.clamp((i0 >> 16) & 0xff, max=max(kz - 1, 1) - 1)
    cell = (bz * max(ky - 1, 1) + by) * max(kx - 1, 1) + bx
    cell = torch.where(row >= 0, cell, torch.full_like(cell, 1 << 20))
    rows = torch.repeat_interleave(torch.arange(n_out, device=dev), rs[1:] - rs[:-1])
    pos_in_row = torch.arange(P, device=dev) - rs[:-1][rows]
    chunk_id = rows * 4096 + pos_in_row // 32
    # runs: positions where the (chunk, cell) key changes
    key = chunk_id * 2048 + torch.clamp(cell, max=2047)
    change = torch.ones(P, dtype=torch.bool, device=dev)
    change[1:] = key[1:] != key[:-1]
    n_runs = int(change.sum())
    valid = int((row >= 0).sum())
    n_chunks = int(torch.unique(chunk_id).numel())
    run_start = torch.nonzero(change).flatten()
    run_len = torch.diff(torch.cat([run_start, torch.tensor([P], device=dev)]))
    run_valid = cell[run_start] < (1 << 20)
    rl = run_len[run_valid].float()
    groups = torch.ceil(rl / 8).sum().item()
    print(f"list n_out {n_out} pairs {P} ({P / max(n_out, 1):.0f} per row): chunks {n_chunks}, runs per chunk {run_valid.sum().item() / n_chunks:.1f}, "
          f"mean run {rl.mean().item():.1f}, groups of 8 per chunk {groups / n_chunks:.1f} (slots used {valid / (8 * groups):.0%})", flush=True)
    return rec


ops.prepare_pair_records = wrapped
import dmcf_b200.models as M
M.ops.prepare_pair_records = wrapped
sim = Simulator(model, device='cuda', step_mode='eager')
with torch.no_grad():
    sim.step(sample)
