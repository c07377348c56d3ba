"""Forward / backward timing of one wide conv layer (32->32, 4x4x4, poly6) at C4 size (1.06 M points, ~35 M pairs)
through dmcf_b200.autograd:  python scripts/bench_backward.py [n_side]"""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from dmcf_b200 import ops, scenes, autograd
n_side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dev = torch.device('cuda')
scene = scenes.lattice_scene((n_side,) * 3, seed=0)
pts = torch.from_numpy(np.concatenate([scene['pos'], scene['box']]).astype(np.float32)).to(dev)
n = pts.shape[0]
g = torch.Generator().manual_seed(0)
W = ((torch.rand((4, 4, 4, 32, 32), generator=g) - 0.5) * 0.1).to(dev).requires_grad_(True)
F = torch.randn((n, 32), generator=g).to(dev).requires_grad_(True)
ext = 0.2
nns = ops.fixed_radius_search(pts, pts, 0.5 * ext, return_distances=False)
print('points', n, 'pairs', nns.neighbors_index.shape[0])
kw = dict(align_corners=True, coordinate_mapping='ball_to_cube_volume_preserving', interpolation='linear', window='poly6', relu_input=True)
d_out = torch.randn((n, 32), generator=g).to(dev)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def fwd():
    with torch.no_grad():
        return ops.continuous_conv(W.detach(), pts, ext, None, pts, F.detach(), None, nns.neighbors_index, None, nns.neighbors_row_splits, **kw)
def fwd_bwd(wgrad=True, fgrad=True):
    W.grad = None; F.grad = None
    W.requires_grad_(wgrad); F.requires_grad_(fgrad)
    out = autograd.continuous_conv(W, pts, ext, pts, F, nns.neighbors_index, nns.neighbors_row_splits, **kw)
    (out * d_out).sum().backward()
t_f = timed(fwd)
t_all = timed(fwd_bwd)
t_w = timed(lambda: fwd_bwd(True, False))
t_x = timed(lambda: fwd_bwd(False, True))
print('forward %.2f ms | forward+backward %.2f ms | fwd + dW only %.2f ms | fwd + dF only %.2f ms' % (t_f, t_all, t_w, t_x))
print('peak memory %.2f GB' % (torch.cuda.max_memory_allocated() / 2**30))
