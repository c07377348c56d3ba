"""Does the shipped Liquid3d checkpoint keep a synthetic block of fluid tame?  (diagnostic)"""
import sys; sys.path.insert(0,'.')
import numpy as np, torch
from dmcf_b200 import config, ops, scenes
from dmcf_b200.simulator import Simulator
z = np.load("tests/golden/ckpt_Liquid3d.npz"); weights = {k.replace("|", "/"): z[k] for k in z.files}
dev=torch.device('cuda')
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
def run(name, sc, grav, steps=60):
    model = config.build_model(dict(scenes.liquid3d_model_cfg(), grav=grav)); model.load_weights(weights, device=dev)
    state=[t(sc['pos']),t(sc['vel']),None,None,t(sc['box']),t(sc['box_normals'])]
    sim=Simulator(model, device='cuda', step_mode='planned')
    out=[]
    with torch.no_grad():
        for i in range(steps):
            state = sim.step(state)
            if i % 10 == 9:
                p, v = state[0], state[1]
                out.append('%d: vmax %.2f vmean %.3f y[%.2f,%.2f] x[%.2f,%.2f]' % (i+1, float(v.norm(dim=1).max()), float(v.norm(dim=1).mean()), float(p[:,1].min()), float(p[:,1].max()), float(p[:,0].min()), float(p[:,0].max())))
    print(name, 'n_box', sc['box'].shape[0], 'replans', sim.stats['replans'], '|', ' | '.join(out), flush=True)
# zero gravity, block away from the walls
sc = scenes.lattice_scene((24,24,24), dx=0.05, jitter=0.1, vel_sigma=0.02, seed=2)
big = scenes.lattice_scene((48,48,48), dx=0.05, seed=1)
sc['pos'] = sc['pos'] + 0.6
sc['box'], sc['box_normals'] = big['box'], big['box_normals']
run('zero-g free block', sc, 0.0)
run('zero-g block in tight box', scenes.lattice_scene((24,24,24), dx=0.05, jitter=0.1, vel_sigma=0.02, seed=2), 0.0)
zc=np.load('tests/golden/canyon_crop.npz')
run('canyon crop g', dict(pos=zc['pos'],vel=zc['vel'],box=zc['box'],box_normals=zc['box_normals']), -9.81, 80)
