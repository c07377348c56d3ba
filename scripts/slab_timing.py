"""Per-step device / host time of the slab step under torchrun, for each step mode (diagnostic).
  torchrun --nproc-per-node N scripts/slab_timing.py [n_side] [steps]"""
import os, sys, time
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from dmcf_b200 import config, ops, scenes
from dmcf_b200.simulator import Simulator
from dmcf_b200.slab import SlabContext

n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
modes = sys.argv[3].split(',') if len(sys.argv) > 3 else ['eager', 'planned']
if os.environ.get('DMCF_OPTIONS'):
    ops.set_kernel_options(int(os.environ['DMCF_OPTIONS']))
world, rank, lr = int(os.environ['WORLD_SIZE']), int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
full = scenes.lattice_scene((n_side,) * 3, dx=0.05, jitter=0.2, vel_sigma=0.1, seed=0)
scene, faces = scenes.slab_partition(full, rank, world, n_side)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
sample = [t(scene['pos']), t(scene['vel']), None, None, t(scene['box']), t(scene['box_normals'])]
for mode in modes:
    model = config.build_model(scenes.c4_model_cfg())
    model.init_weights(seed=0, device=dev, scale=0.1)
    model.set_slab(SlabContext(faces, axis=0))
    sim = Simulator(model, device=f'cuda:{lr}', step_mode=mode)
    dev_ms, host_ms = [], []
    with torch.no_grad():
        for i in range(3 + steps):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time(); e0.record()
            sim.step(sample)
            e1.record(); t1 = time.time()
            torch.cuda.synchronize()
            if i >= 3:
                dev_ms.append(e0.elapsed_time(e1)); host_ms.append((t1 - t0) * 1e3)
    with torch.no_grad():
        ops.PROFILE = []
        for _ in range(3):
            sim.step(sample)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    if rank == 0:
        g = {}
        for r in prof:
            k = r.get('kind') or (r['kernel'], r['cin'], r['cout'], r['n_out'], r['pairs'])
            g.setdefault(k, []).append(r['start'].elapsed_time(r['end']))
        for k, v in g.items():
            print('   ', k, 'avg ms %.3f x %d per step' % (sum(v) / len(v), len(v) // 3))
    if rank == 0:
        print(f'{mode:8s} world {world} n_own {scene["pos"].shape[0]}: device ms/step median {np.median(dev_ms):.2f} '
              f'(min {min(dev_ms):.2f} max {max(dev_ms):.2f}), host enqueue ms/step median {np.median(host_ms):.2f}, stats {sim.stats}', flush=True)
    if mode == 'planned' and len(sys.argv) > 4:  # torch profiler summary of three steps
        from torch.profiler import profile, ProfilerActivity
        with torch.no_grad(), profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                sim.step(sample)
            torch.cuda.synchronize()
        if rank == 0:
            print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25))
            print(prof.key_averages().table(sort_by='cpu_time_total', row_limit=25))
dist.barrier(); torch.cuda.synchronize(); sys.stdout.flush(); os._exit(0)
