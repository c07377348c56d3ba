"""The continuous_conv parity cases and their seeded inputs, shared by tests/test_ops_gpu.py (CUDA vs oracle), tests/test_oracle_cpu.py
(oracle vs Open3D golden outputs when tests/golden/open3d_conv_cases.npz exists) and scripts/make_open3d_golden.py (which writes
that file on a machine where open3d.ml is importable)."""
import zlib

import numpy as np

CONV_CASES = [
    # name, kernel_size, cin, cout, mapping, interp, align, normalize, window, ignore_q
    ("wide444", (4, 4, 4), 32, 32, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False),
    ("inp444", (4, 4, 4), 4, 8, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False),
    ("c24", (4, 4, 4), 24, 32, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False),
    ("k188", (1, 8, 8), 7, 8, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False),
    ("k181", (1, 8, 1), 8, 16, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", True),
    ("radial_norm", (3, 3, 3), 5, 6, "ball_to_cube_radial", "linear", True, True, None, False),
    ("radial_norm_win", (3, 3, 3), 5, 6, "ball_to_cube_radial", "linear", True, True, "cubic", False),
    ("identity_noalign", (4, 3, 2), 3, 2, "identity", "linear", False, False, "linear", False),
    ("border", (3, 3, 3), 4, 4, "ball_to_cube_radial", "linear_border", False, False, "peak", True),
    ("nearest", (3, 3, 3), 4, 4, "ball_to_cube_volume_preserving", "nearest_neighbor", True, False, None, False),
    ("cout64", (4, 4, 4), 16, 64, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", True),
    ("cin96", (4, 4, 4), 96, 64, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", True),
    ("cout3", (6, 6, 6), 32, 3, "ball_to_cube_volume_preserving", "linear", True, False, "peak", True),
    ("sampling111", (1, 1, 1), 3, 3, "ball_to_cube_radial", "linear", True, True, "poly6", False),
    ("cubic_grad", (2, 2, 2), 2, 1, "ball_to_cube_radial", "linear", True, False, "cubic_grad", False),
    ("wide188", (1, 8, 8), 32, 32, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False),
    ("wide181", (1, 8, 1), 24, 16, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", True),
    ("wide444_norm_radial", (4, 4, 4), 20, 40, "ball_to_cube_radial", "linear", False, True, "cubic", False),
    ("direct_cout1_cin40", (3, 3, 3), 40, 1, "ball_to_cube_radial", "linear", True, True, "poly6", False),
    ("direct_cout2_188", (1, 8, 8), 32, 2, "ball_to_cube_volume_preserving", "linear", True, False, "peak", True),
    ("direct_cout4_border", (3, 3, 3), 6, 4, "ball_to_cube_radial", "linear_border", False, False, "peak", True),
    # long neighbour rows (mean ~60, max > 96: several 32-pair chunks per out point, chunk prefetch across points)
    ("wide444_long_rows", (4, 4, 4), 32, 32, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False, 0.5),
    ("wide188_long_rows", (1, 8, 8), 16, 24, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", True, 0.3),
    ("wide444_cin7_cout12", (4, 4, 4), 7, 12, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False, 0.4),
    # lattice without jitter: neighbours exactly on the filter border (g == fs-1), ties in the cell order
    ("wide444_lattice", (4, 4, 4), 8, 8, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", False, "lattice"),
    # tensor-core phase 2 of k_cconv_lean (cout == 32, cin % 8 == 0) with the normaliser, and on the thin grids
    ("tc444_norm", (4, 4, 4), 16, 32, "ball_to_cube_radial", "linear", True, True, "cubic", False),
    ("tc181", (1, 8, 1), 16, 32, "ball_to_cube_volume_preserving", "linear", True, False, "poly6", True),
    ("tc444_long_rows_cin24", (4, 4, 4), 24, 32, "ball_to_cube_volume_preserving", "linear", True, False, "peak", True, 0.45),
]


def conv_case_inputs(case):
    """(ks, cin, cout, mapping, interp, align, normalize, window, ignore_q, pts, outp, feats, filt, extent, radius) of one case."""
    name, ks, cin, cout, mapping, interp, align, normalize, window, ignore_q = case[:10]
    variant = case[10] if len(case) > 10 else None
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    n_in, n_out = 900, 700
    pts = rng.random((n_in, 3)).astype(np.float32)
    if ks[0] == 1:
        pts[:, 2] = 0
    if ks[2] == 1:
        pts[:, 0] = 0
    outp = pts[:n_out].copy()
    outp[n_out // 2:] += rng.normal(0, 0.01, (n_out - n_out // 2, 3)).astype(np.float32) * (pts[:1] * 0 + (np.array(ks[::-1]) > 1))
    outp = outp.astype(np.float32)
    if variant == "lattice":  # 10x10x9 lattice of pitch 1/16, extent = 4 pitches: neighbours at exactly r along the axes
        g = np.stack(np.meshgrid(np.arange(10), np.arange(10), np.arange(9), indexing="ij"), -1).reshape(-1, 3)
        pts = (g.astype(np.float32) * np.float32(0.0625))
        outp = pts[:n_out].copy()
    feats = rng.standard_normal((n_in, cin)).astype(np.float32)
    filt = rng.uniform(-0.5, 0.5, ks + (cin, cout)).astype(np.float32)
    extent = np.float32(variant if isinstance(variant, float) else 0.25)
    radius = np.float32(0.5) * extent
    return ks, cin, cout, mapping, interp, align, normalize, window, ignore_q, pts, outp, feats, filt, extent, radius
