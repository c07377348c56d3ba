import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, torch
from dmcf_b200 import config, ops, scenes
from dmcf_b200.simulator import Simulator
z = np.load("tests/golden/ckpt_Liquid3d.npz"); weights = {k.replace("|", "/"): z[k] for k in z.files}
dev=torch.device('cuda')
model = config.build_model(scenes.liquid3d_model_cfg()); model.load_weights(weights, device=dev)
sc = scenes.lattice_scene((40,40,40), dx=0.05, jitter=0.2, vel_sigma=0.05, seed=2, open_top=True)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
state=[t(sc['pos']),t(sc['vel']),None,None,t(sc['box']),t(sc['box_normals'])]
sim=Simulator(model, device='cuda', step_mode=sys.argv[1] if len(sys.argv)>1 else 'planned')
with torch.no_grad():
    for i in range(10):
        # what would a measuring step record now?
        plan = ops.StepPlan(dev); plan.begin('measure'); ops.set_plan(plan); model(state); ops.set_plan(None)
        lat=[e for e in plan.entries if e['kind']=='lattice']
        cur = sim._planned.plan.entries if sim._planned and sim._planned.plan else None
        print(i, 'now', [(e['lo'],e['dims'],e['count']) for e in lat], 'plan', [(e['lo'],e['dims'],e['count']) for e in cur if e['kind']=='lattice'] if cur else None)
        state = sim.step(state)
        print('   stats', sim.stats, sim.replan_log[-1:] )
