"""A/B timing of the wide-layer kernels on the C4 scene: per-kernel CUDA-event times of one planned step for a list of
kernel-option masks (dmcf_set_kernel_options): python scripts/ws_ab.py 3 131 259 2051 ..."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from dmcf_b200 import ops, config, scenes
from dmcf_b200.simulator import Simulator
dev = torch.device('cuda')
n = 100
scene = scenes.lattice_scene((n, n, n), seed=0)
model = config.build_model(scenes.c4_model_cfg()); model.init_weights(seed=0, device=dev, scale=0.1)
t = lambda a: torch.from_numpy(a).to(dev)
sample = [t(scene['pos']), t(scene['vel']), None, None, t(scene['box']), t(scene['box_normals'])]
for opt in [int(a) for a in sys.argv[1:]] or [3, 131]:
    ops.set_kernel_options(opt)
    sim = Simulator(model, device='cuda', step_mode='planned')
    with torch.no_grad():
        for _ in range(3):
            sim.step(sample)
        ops.PROFILE = []
        torch.cuda.synchronize()
        for _ in range(3):
            sim.step(sample)
        torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    g = {}
    for r in prof:
        if 'kind' in r:
            continue
        g.setdefault((r['kernel'], r['cin'], r['cout']), []).append(r['start'].elapsed_time(r['end']))
    print('options', opt, ' '.join('%s %d->%d %.3f ms' % (k[0], k[1], k[2], sum(v) / len(v)) for k, v in g.items()), flush=True)
