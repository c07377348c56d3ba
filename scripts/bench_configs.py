"""Per-step timing of the BASELINE configs 2 and 3 (WBC-SPH 2-D ~10k particles, Liquid3d 3-D ~100k particles) with the
shipped checkpoints (tests/golden/ckpt_*.npz): ms/step and the time per conv kernel group.
`c5shard` is one GPU's share of BASELINE config 5 (4 M particles on 8 GPUs, full multi-scale Liquid3d net): 80^3 = 512 000
fluid particles in an open box.
  python scripts/bench_configs.py [liquid3d|wbc|c5shard|c4] [steps] [kernel options]"""
import os, sys
sys.path.insert(0, '.')
sys.path.insert(0, 'tests')
import numpy as np, torch
from dmcf_b200 import ops, config, scenes
from dmcf_b200.simulator import Simulator
import test_models_gpu as T
which = sys.argv[1] if len(sys.argv) > 1 else 'liquid3d'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
if len(sys.argv) > 3:
    ops.set_kernel_options(int(sys.argv[3]))  # e.g. 35 = single-pair walk for narrow inputs (A/B against the default 3)
dev = torch.device('cuda')
if which == 'c4':
    cfg, scene, wname = scenes.c4_model_cfg(), scenes.lattice_scene((100, 100, 100), dx=0.05, seed=0), None
    acc = None
elif which == 'c5shard':
    cfg, scene, wname = T.liquid3d_cfg(), scenes.lattice_scene((80, 80, 80), dx=0.05, seed=2, open_top=True), 'ckpt_Liquid3d.npz'
    acc = None
elif which == 'liquid3d':
    cfg, scene, wname = T.liquid3d_cfg(), scenes.lattice_scene((46, 46, 46), dx=0.05, seed=2, open_top=True), 'ckpt_Liquid3d.npz'
    acc = None
else:
    cfg, scene, wname = T.wbc_cfg(), scenes.lattice_scene((100, 100, 1), dx=0.005, seed=3, vel_sigma=0.05), 'ckpt_WBC-SPH.npz'
    acc = np.tile(np.array([[0.0, -9.81, 0.0]], np.float32), (scene['pos'].shape[0], 1))
model = config.build_model(cfg)
if wname is None:
    model.init_weights(seed=0, device=dev, scale=0.1)
else:
    model.load_weights(T.load_npz_weights(wname), device=dev)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
sample = [t(scene['pos']), t(scene['vel']), None if acc is None else t(acc), None, t(scene['box']), t(scene['box_normals'])]
print(which, 'fluid', scene['pos'].shape[0], 'boundary', scene['box'].shape[0])


def timed(mode, steps, profile=False):
    sim = Simulator(model, device='cuda', step_mode=mode)
    with torch.no_grad():
        for _ in range(4): sim.step(sample)
        ops.PROFILE = [] if profile else None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launch_count()
        e0.record()
        for _ in range(steps): out = sim.step(sample)
        e1.record(); torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    return e0.elapsed_time(e1) / steps, prof, (ops.launch_count() - l0) // steps, sim


for mode in ('eager', 'planned', 'graph'):
    ms, _, launches, sim = timed(mode, steps)
    print('%-8s ms/step %.3f  particles*steps/s %.3e  (host launches/step %d, graph kernels/step %d, stats %s)' % (
        mode, ms, scene['pos'].shape[0] / ms * 1e3, launches, sim._planned.graph_launches if sim._planned else 0, sim.stats))
ms, prof, _, _ = timed('planned', steps, profile=True)
hbm = [r for r in prof if 'kind' in r]
prof = [r for r in prof if 'kind' not in r]  # conv launches only
print('per-kernel times below: planned mode with per-launch events, ms/step %.2f' % ms)
g = {}
for r in prof:
    k = (r['kernel'], r['kernel_size'], r['cin'], r['cout'], r['n_inp'], r['n_out'], r['pairs'])
    g.setdefault(k, []).append(r['start'].elapsed_time(r['end']))
tot = 0
for k, v in sorted(g.items(), key=lambda kv: -sum(kv[1])):
    per = sum(v) / steps
    tot += per
    print('%-15s %s %3d->%-3d n_in %7d n_out %7d pairs %9d : %6.3f ms/step (%d launches/step)' % (k[0], k[1], k[2], k[3], k[4], k[5], k[6], per, len(v) // steps))
print('conv kernels total %.2f ms/step' % tot)
h = {}
for r in hbm:
    a = h.setdefault(r['kind'], [0.0, 0, 0])
    a[0] += r['start'].elapsed_time(r['end']); a[1] += 1; a[2] += r.get('pairs', 0)
for k, a in sorted(h.items(), key=lambda kv: -kv[1][0]):
    print('%-13s %7.3f ms/step (%d launches/step, %d pairs/step)' % (k, a[0] / steps, a[1] // steps, a[2] // steps))
print('peak memory %.2f GB' % (torch.cuda.max_memory_allocated() / 2 ** 30))
