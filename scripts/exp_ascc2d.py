"""A/B of the antisymmetric 2-D layer (1x8x8, 32->2): k_cconv_apatch vs k_cconv_direct."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from dmcf_b200 import ops
dev = torch.device('cuda')
rng = np.random.default_rng(0)
n = 400 * 400
g = np.stack(np.meshgrid(np.arange(400), np.arange(400), indexing='ij'), -1).reshape(-1, 2).astype(np.float32)
pts = np.zeros((n, 3), np.float32); pts[:, :2] = (g + rng.uniform(-0.2, 0.2, g.shape)) * 0.005
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
p = t(pts); feats = t(rng.standard_normal((n, 32)).astype(np.float32))
half = rng.uniform(-0.5, 0.5, (1, 4, 8, 32, 2)).astype(np.float32)
full = t(np.concatenate([-half[::-1, ::-1, ::-1], half], axis=1))
ext = 0.02
nns = ops.fixed_radius_search(p, p, 0.5 * ext, ignore_query_point=True)
print('pairs/point', nns.neighbors_index.shape[0] / n)
recs = ops.prepare_pair_records((1, 8, 8), p, ext, None, p, None, nns.neighbors_index, None, nns.neighbors_row_splits,
                                align_corners=True, coordinate_mapping='ball_to_cube_volume_preserving', interpolation='linear', window='peak')
def run(flag):
    f = lambda: ops.continuous_conv(full, p, ext, None, p, feats, None, nns.neighbors_index, None, nns.neighbors_row_splits,
                                    align_corners=True, coordinate_mapping='ball_to_cube_volume_preserving', normalize=False,
                                    interpolation='linear', window='peak', relu_input=True, ascc=True,
                                    antisymmetric_filter=flag, pair_records=recs)
    for _ in range(3): out = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): out = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10, out
ta, oa = run(True); tb, ob = run(False)
print('apatch %.3f ms  direct %.3f ms  max diff %.2e' % (ta, tb, float((oa - ob).abs().max())))
