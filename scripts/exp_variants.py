import sys, json, subprocess
sys.path.insert(0, '.')
import torch, numpy as np
from dmcf_b200 import ops, config, scenes
from dmcf_b200.simulator import Simulator
dev = torch.device('cuda')
scene = scenes.lattice_scene((100,100,100), seed=0)
model = config.build_model(scenes.c4_model_cfg()); model.init_weights(seed=0, device=dev, scale=0.1)
sim = Simulator(model, device='cuda')
t = lambda a: torch.from_numpy(a).to(dev)
sample = [t(scene['pos']), t(scene['vel']), None, None, t(scene['box']), t(scene['box_normals'])]
def run(opt, label):
    ops.set_kernel_options(opt)
    for _ in range(2): sim.step(sample)
    ops.PROFILE = []
    torch.cuda.synchronize(); 
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): sim.step(sample)
    e1.record(); torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    prof = [r for r in prof if 'kind' not in r]
    g = {}
    for r in prof:
        k = (r['kernel_size'], r['cin'], r['cout'])
        g.setdefault(k, []).append(r['start'].elapsed_time(r['end']))
    print(label, 'ms/step %.2f' % (e0.elapsed_time(e1)/3), {str(k): round(float(np.mean(v)),2) for k,v in g.items()})
run(3, 'lean + apatch')
run(27, 'legacy wide + direct')
run(3 | 512, 'lean, phase 1 only')
