"""Runs a few C4 steps (100^3 fluid lattice, single-scale SymNet) -- the workload ncu is wrapped around:
  ncu --set full --clock-control none --import-source on -k regex:k_cconv_lean -s 8 -c 3 -o gpurun_out/prof \
      python scripts/run_steps.py 3 [kernel options]"""
import sys
sys.path.insert(0, '.')
import torch
from dmcf_b200 import ops, config, scenes
from dmcf_b200.simulator import Simulator
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
if len(sys.argv) > 2:
    ops.set_kernel_options(int(sys.argv[2]))
n = int(sys.argv[3]) if len(sys.argv) > 3 else 100
dev = torch.device('cuda')
scene = scenes.lattice_scene((n, n, n), seed=0)
model = config.build_model(scenes.c4_model_cfg()); model.init_weights(seed=0, device=dev, scale=0.1)
sim = Simulator(model, device='cuda')
t = lambda a: torch.from_numpy(a).to(dev)
sample = [t(scene['pos']), t(scene['vel']), None, None, t(scene['box']), t(scene['box_normals'])]
for _ in range(steps):
    sim.step(sample)
torch.cuda.synchronize()
print("done", steps)
