"""Small launches of the round-2 kernels for compute-sanitizer (memcheck / racecheck):
  compute-sanitizer --tool racecheck python scripts/sanitize_kernels.py
k_cconv_lean with the tensor-core phase 2 as persistent CTAs (more tiles than SMs), k_cconv_narrow, k_cconv_direct with in-kernel
geometry, k_dense_umma."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
from dmcf_b200 import ops
dev = torch.device('cuda')
rng = np.random.default_rng(0)
n = 4200  # 175 tiles of 24 points: the persistent loop runs more than once on some CTAs
pts = torch.from_numpy((rng.random((n, 3)) * 1.2).astype(np.float32)).to(dev)
nns = ops.fixed_radius_search(pts, pts, 0.1)
kw = dict(align_corners=True, coordinate_mapping="ball_to_cube_volume_preserving", normalize=False, interpolation="linear", window="poly6")
f32 = torch.randn((n, 32), device=dev)
w = torch.randn((4, 4, 4, 32, 32), device=dev) * 0.1
wd = torch.randn((32, 32), device=dev) * 0.1
w_ext = torch.cat([w.reshape(-1, 32), wd], dim=0).contiguous()
out = ops.continuous_conv(w_ext, pts, 0.2, None, pts, f32, None, nns.neighbors_index, None, nns.neighbors_row_splits, relu_input=True,
                          dense_inp=f32, dense_cin=32, kernel_size=(4, 4, 4), residual=f32, **kw)
x8 = torch.zeros((n, 8), device=dev); x8[: n // 2, :4] = torch.randn((n // 2, 4), device=dev); x8[n // 2:, 4:] = torch.randn((n - n // 2, 4), device=dev)
wb = torch.zeros((64, 8, 24), device=dev); wb[:, :4, :8] = torch.randn((64, 4, 8), device=dev); wb[:, 4:, 8:16] = torch.randn((64, 4, 8), device=dev)
wb_ext = torch.cat([wb.reshape(-1, 24), torch.randn((8, 24), device=dev)], dim=0).contiguous()
out2 = ops.continuous_conv(wb_ext, pts, 0.2, None, pts, x8, None, nns.neighbors_index, None, nns.neighbors_row_splits, dense_inp=x8,
                           dense_cin=8, kernel_size=(4, 4, 4), block_diagonal=(4, 8, 8), **kw)
half = torch.randn((6, 3, 6, 32, 3), device=dev) * 0.1
full = torch.cat([-torch.flip(half, dims=(0, 1, 2)), half], dim=1).contiguous()
out3 = ops.continuous_conv(full, pts, 0.2, None, pts, f32, None, nns.neighbors_index, None, nns.neighbors_row_splits, ascc=True,
                           skip_self=True, relu_input=True, antisymmetric_filter=True, **kw)
out4 = ops.dense(torch.randn((5000, 32), device=dev), wd, torch.randn(32, device=dev), relu_input=True)
torch.cuda.synchronize()
print("ok", float(out.abs().sum()), float(out2.abs().sum()), float(out3.abs().sum()), float(out4.abs().sum()))
