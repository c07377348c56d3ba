"""GPU parity tests proper: the CUDA path (through the C ABI) against the O64 oracle on identical seeded inputs.

Bars: neighbour rows bit-exact after a per-row sort (indices AND squared distances);
features within 2e-5*max|ref| + 1e-6 of the float64 oracle.
"""
import zlib

import numpy as np
import pytest
import torch

from oracle import o64

pytestmark = pytest.mark.gpu

RTOL_MAX, ATOL = 2e-5, 1e-6


def feat_close(got, ref, scale=1.0):
    ref = np.asarray(ref, np.float64)
    got = np.asarray(got, np.float64)
    tol = scale * (RTOL_MAX * max(np.abs(ref).max(initial=0.0), 1e-30) + ATOL)
    err = np.abs(got - ref).max(initial=0.0)
    assert err <= tol, f"max err {err:.3e} > tol {tol:.3e} (max|ref| {np.abs(ref).max(initial=0.0):.3e})"


def sorted_rows(index, splits, dist=None):
    index = np.asarray(index)
    splits = np.asarray(splits)
    rows = np.repeat(np.arange(len(splits) - 1), np.diff(splits))
    order = np.lexsort((index, rows))
    return index[order], (None if dist is None else np.asarray(dist)[order])


def clouds():
    rng = np.random.default_rng(7)
    lattice = np.stack(np.meshgrid(*[np.arange(12) * 0.05] * 3, indexing="ij"), -1).reshape(-1, 3)
    yield "uniform3d", rng.random((4000, 3)).astype(np.float32), 0.1
    yield "jitter_lattice", (lattice + rng.uniform(-0.01, 0.01, lattice.shape)).astype(np.float32), 0.1
    yield "exact_lattice_ties", lattice.astype(np.float32), 0.1  # neighbours at exactly 2 spacings (float ties)
    p2 = rng.random((3000, 3)).astype(np.float32); p2[:, 2] = 0
    yield "planar2d", p2, 0.04
    p1 = np.zeros((500, 3), np.float32); p1[:, 1] = np.sort(rng.random(500)) * 2
    yield "column1d", p1, 0.02
    dup = rng.random((300, 3)).astype(np.float32)
    yield "duplicates", np.concatenate([dup, dup[:100]]), 0.15
    yield "single", np.array([[0.3, 0.2, 0.1]], np.float32), 0.5
    far = rng.random((1000, 3)).astype(np.float32) * np.array([100, 1, 1], np.float32) + 1000.0
    yield "large_coords", far.astype(np.float32), 0.3


@pytest.fixture(params=[3, 67], ids=["cell-centric-search", "query-centric-search"])
def frs_options(request):
    """Same-set searches (queries = a prefix of the points) run k_frs_cell by default; option bit 6 keeps k_frs."""
    from dmcf_b200 import ops
    prev = ops.set_kernel_options(request.param)
    yield request.param
    ops.set_kernel_options(prev)


@pytest.mark.parametrize("name,pts,radius", list(clouds()), ids=[c[0] for c in clouds()])
@pytest.mark.parametrize("ignore", [False, True])
def test_fixed_radius_search_bit_exact(cuda, name, pts, radius, ignore, frs_options):
    from dmcf_b200 import ops
    t = torch.from_numpy(pts).to(cuda)
    res = ops.fixed_radius_search(t, t, radius, ignore_query_point=ignore, return_distances=True)
    ri, rs, rd = o64.fixed_radius_search(pts, pts, radius, ignore_query_point=ignore)
    gi, gs, gd = res.neighbors_index.cpu().numpy(), res.neighbors_row_splits.cpu().numpy(), res.neighbors_distance.cpu().numpy()
    assert gi.dtype == np.int32 and gs.dtype == np.int64 and gd.dtype == np.float32
    assert np.array_equal(gs, rs)
    a_i, a_d = sorted_rows(gi, gs, gd)
    b_i, b_d = sorted_rows(ri, rs, rd)
    assert np.array_equal(a_i, b_i)
    assert np.array_equal(a_d, b_d)  # squared distances bit-exact


def test_fixed_radius_search_distinct_sets_and_empty(cuda):
    from dmcf_b200 import ops
    rng = np.random.default_rng(3)
    pts = rng.random((2500, 3)).astype(np.float32)
    qs = (rng.random((700, 3)) * 1.4 - 0.2).astype(np.float32)  # some queries outside the cell grid
    res = ops.fixed_radius_search(torch.from_numpy(pts).to(cuda), torch.from_numpy(qs).to(cuda), 0.08)
    ri, rs, rd = o64.fixed_radius_search(pts, qs, 0.08)
    assert np.array_equal(res.neighbors_row_splits.cpu().numpy(), rs)
    a_i, a_d = sorted_rows(res.neighbors_index.cpu().numpy(), rs, res.neighbors_distance.cpu().numpy())
    b_i, b_d = sorted_rows(ri, rs, rd)
    assert np.array_equal(a_i, b_i) and np.array_equal(a_d, b_d)
    # empty point set / empty query set
    e = torch.zeros((0, 3), device=cuda)
    r0 = ops.fixed_radius_search(e, torch.from_numpy(qs).to(cuda), 0.1)
    assert r0.neighbors_index.numel() == 0 and int(r0.neighbors_row_splits.abs().sum()) == 0
    r1 = ops.fixed_radius_search(torch.from_numpy(pts).to(cuda), e, 0.1)
    assert r1.neighbors_row_splits.shape[0] == 1 and r1.neighbors_index.numel() == 0
    # deterministic row order: two builds give identical arrays
    t = torch.from_numpy(pts).to(cuda)
    x = ops.fixed_radius_search(t, t, 0.1)
    y = ops.fixed_radius_search(t, t, 0.1)
    assert torch.equal(x.neighbors_index, y.neighbors_index)


def test_prefix_search_kernels_agree_row_for_row(cuda):
    """The cell-centric kernel (queries = the first rows of the point set, as in every same-set search of a step: owned rows
    out, owned + ghost rows in) returns the SAME arrays as the query-centric one, element for element (row order included),
    also for a proper prefix, with crowded cells (> 32 points per cell) and with a device-side query count."""
    from dmcf_b200 import ops
    rng = np.random.default_rng(9)
    pts = np.concatenate([rng.random((6000, 3)), rng.random((3000, 3)) * 0.05 + 0.4]).astype(np.float32)  # a crowded clump
    pts = pts[rng.permutation(len(pts))]
    t = torch.from_numpy(pts).to(cuda)
    for nq, radius in ((len(pts), 0.06), (5000, 0.06), (1, 0.2), (7000, 0.11)):
        q = t[:nq]  # a view: same base pointer -> prefix case
        out = {}
        for opt in (3, 67):
            prev = ops.set_kernel_options(opt)
            try:
                cl = ops.CellList(t, radius)
                out[opt] = ops.fixed_radius_search(t, q, radius, cell_list=cl, return_distances=True)
            finally:
                ops.set_kernel_options(prev)
        for a, b in zip(out[3], out[67]):
            assert torch.equal(a, b), (nq, radius)
        ri, rs, rd = o64.fixed_radius_search(pts, pts[:nq], radius)
        assert np.array_equal(out[3].neighbors_row_splits.cpu().numpy(), rs)
    # device-side count: rows beyond it are empty
    cl = ops.CellList(t, 0.06)
    nq_dev = torch.tensor([4000], dtype=torch.int32, device=cuda)
    res = ops.fixed_radius_search(t, ops.with_count(t[:6000], nq_dev), 0.06, cell_list=cl, capacity=8_000_000)
    ref = ops.fixed_radius_search(t, t[:4000], 0.06, cell_list=cl)
    assert torch.equal(res.neighbors_row_splits[:4001], ref.neighbors_row_splits)
    assert bool((res.neighbors_row_splits[4001:] == ref.neighbors_row_splits[-1]).all())
    n = int(ref.neighbors_row_splits[-1])
    assert torch.equal(res.neighbors_index[:n], ref.neighbors_index)


def test_neighbor_counts_is_reduce_subarrays_sum(cuda):
    from dmcf_b200 import ops
    rng = np.random.default_rng(5)
    pts = rng.random((3000, 3)).astype(np.float32)
    t = torch.from_numpy(pts).to(cuda)
    counts, _ = ops.neighbor_counts(t, t, 0.09)
    _, rs, _ = o64.fixed_radius_search(pts, pts, 0.09)
    assert np.array_equal(counts.cpu().numpy(), np.diff(rs).astype(np.int32))


def test_exclusive_scan_large(cuda):
    from dmcf_b200 import ops
    rng = np.random.default_rng(1)
    for n in (0, 1, 5, 2048, 2049, 100_000, 3_000_001):
        a = rng.integers(0, 2000, n).astype(np.int32)
        out = ops.exclusive_scan(torch.from_numpy(a).to(cuda), torch.int64).cpu().numpy()
        ref = np.concatenate([[0], np.cumsum(a.astype(np.int64))])
        assert np.array_equal(out, ref), n


from conv_cases import CONV_CASES, conv_case_inputs  # noqa: E402


@pytest.fixture(params=[3, 35, 131, 27, 31, 0, 3 | 32768, 3 | 65536],
                ids=["fast-kernels", "fast-kernels-single-pair-walk", "warp-specialised-experiment", "legacy-wide-direct",
                     "legacy-wide-zsplit-direct", "generic-kernel", "fast-kernels-ffma2-phase2", "fast-kernels-tc-16x16"])
def kernel_options(request):
    from dmcf_b200 import ops
    prev = ops.set_kernel_options(request.param)
    yield request.param
    ops.set_kernel_options(prev)


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
@pytest.mark.parametrize("fused_window", [False, True])
def test_continuous_conv_matches_oracle(cuda, case, fused_window, kernel_options):
    from dmcf_b200 import ops
    ks, cin, cout, mapping, interp, align, normalize, window, ignore_q, pts, outp, feats, filt, extent, radius = conv_case_inputs(case)
    idx, splits, d2 = o64.fixed_radius_search(pts, outp, radius, ignore_query_point=ignore_q)
    imp = None
    if window is not None:
        imp = o64.window(window, d2.astype(np.float64) / (np.float64(radius) ** 2))
    ref = o64.continuous_conv(filt, outp, extent, (0, 0, 0), pts, feats, None, idx, imp, splits, align_corners=align,
                              coordinate_mapping=mapping, normalize=normalize, interpolation=interp)
    dev = cuda
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    kw = dict(align_corners=align, coordinate_mapping=mapping, normalize=normalize, interpolation=interp)
    if fused_window or window is None:
        got = ops.continuous_conv(t(filt), t(outp), float(extent), None, t(pts), t(feats), None, t(idx), None, t(splits),
                                  window=window, **kw)
    else:
        got = ops.continuous_conv(t(filt), t(outp), float(extent), None, t(pts), t(feats), None, t(idx),
                                  t(imp.astype(np.float32)), t(splits), **kw)
    scale = 4.0 if cin * np.prod(ks) > 4096 else 1.0  # very long float32 dot products (cin96: K = 6144)
    feat_close(got.cpu().numpy(), ref, scale)
    if not normalize and (fused_window or window is None):
        # same conv through precomputed pair records (dmcf_cconv_prepare; pairs reordered by filter cell)
        recs = ops.prepare_pair_records(ks, t(outp), float(extent), None, t(pts), None, t(idx), None, t(splits),
                                        align_corners=align, coordinate_mapping=mapping, interpolation=interp, window=window)
        got2 = ops.continuous_conv(t(filt), t(outp), float(extent), None, t(pts), t(feats), None, t(idx), None, t(splits),
                                   window=window, pair_records=recs, **kw)
        feat_close(got2.cpu().numpy(), ref, scale)
        feat_close(got2.cpu().numpy(), got.cpu().numpy(), 0.5)


@pytest.mark.parametrize("cin", [24, 16, 11, 8, 3])
def test_continuous_conv_fused_extras(cuda, kernel_options, cin):
    """relu on the input, feature scale, neighbour sub-range, skip-self on a self-containing CSR, fused Dense,
    bias, residual, accumulate, strided in/out rows.  cin <= 16 runs the multi-pair phase 1 of k_cconv_lean (2 / 4 / 8 pair
    slots per step) unless the kernel options ask for the single-pair walk."""
    from dmcf_b200 import ops
    rng = np.random.default_rng(11)
    n, cout, ks = 800, 32, (4, 4, 4)
    pts = rng.random((n, 3)).astype(np.float32)
    feats_wide = rng.standard_normal((n, cin + 5)).astype(np.float32)
    feats = feats_wide[:, 2:2 + cin]
    filt = rng.uniform(-0.3, 0.3, ks + (cin, cout)).astype(np.float32)
    dk = rng.uniform(-0.3, 0.3, (cin, cout)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    resid = rng.standard_normal((n, cout)).astype(np.float32)
    extent = np.float32(0.3)
    radius = np.float32(0.5) * extent
    idx, splits, d2 = o64.fixed_radius_search(pts, pts, radius, ignore_query_point=False)  # contains self
    lo, hi = 100, 650
    # reference: relu, scale, ignore self, only neighbours in [lo,hi)
    keep = (idx >= lo) & (idx < hi) & ~np.all(pts[idx] == np.repeat(pts, np.diff(splits), axis=0), axis=1)
    rows = np.repeat(np.arange(n), np.diff(splits))[keep]
    k_idx = idx[keep]
    k_splits = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int64)
    imp = o64.window("poly6", d2[keep].astype(np.float64) / np.float64(radius) ** 2)
    g = np.maximum(feats, 0) * 0.5
    ref = o64.continuous_conv(filt, pts, extent, (0, 0, 0), pts, g, None, k_idx, imp, k_splits, align_corners=True,
                              coordinate_mapping="ball_to_cube_volume_preserving", normalize=False, interpolation="linear")
    ref = ref + np.maximum(feats, 0).astype(np.float64) @ dk + bias + resid
    prev = rng.standard_normal((n, cout)).astype(np.float32)
    ref_acc = ref + prev
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    fw = t(feats_wide)
    w_ext = torch.cat([t(filt).reshape(-1, cout), t(dk)], dim=0)
    out_wide = torch.zeros((n, cout + 7), device=cuda)
    out_view = out_wide[:, 3:3 + cout]
    out_view.copy_(t(prev))
    # neighbour sub-range: features/positions of the sub-range are passed as their own arrays (row = index - lo)
    ops.continuous_conv(w_ext, t(pts), float(extent), None, t(pts[lo:hi]), fw[lo:hi, 2:2 + cin], None, t(idx), None, t(splits),
                        align_corners=True, coordinate_mapping="ball_to_cube_volume_preserving", normalize=False,
                        interpolation="linear", window="poly6", relu_input=True, feat_scale=0.5, skip_self=True,
                        nbr_range=(lo, hi), bias=t(bias), dense_inp=fw[:, 2:2 + cin], dense_cin=cin, residual=t(resid),
                        out=out_view, accumulate=True, kernel_size=ks)
    feat_close(out_view.cpu().numpy(), ref_acc)
    assert float(out_wide[:, :3].abs().sum()) == 0 and float(out_wide[:, 3 + cout:].abs().sum()) == 0


@pytest.mark.parametrize("shape", [(4, 4, 8, 8, 8), (3, 2, 5, 7, 4), (1, 4, 8, 3, 0), (4, 1, 2, 8, 8)],
                         ids=["c4-input-layer", "ragged-groups", "one-channel-group", "one-box-channel"])
@pytest.mark.parametrize("with_records", [False, True])
def test_block_diagonal_input_layer(cuda, shape, with_records):
    """The fused input layer of the DMCF nets (models/pbf_model.py:375-411: fluid_convs + obs_convs + two Dense over zero-padded
    [fluid | box] rows, one block diagonal filter): the narrow direct kernel (k_cconv_narrow, takes the block promise of
    dmcf_conv_desc::block_cin) against the O64 oracle of the zero-padded conv, and against the register-patch route (option bit 12)
    which ignores the promise.  Rows of the two kinds are interleaved (ghost rows arrive in any order)."""
    from dmcf_b200 import ops
    ca, cb, na, nb, n_dense = shape
    rng = np.random.default_rng(sum(shape))
    n, ks = 900, (4, 4, 4)
    cin, cout = ca + cb, na + nb + n_dense
    pts = rng.random((n, 3)).astype(np.float32)
    kind_b = rng.random(n) < 0.3
    feats = rng.standard_normal((n, cin)).astype(np.float32)
    feats[kind_b, :ca] = 0
    feats[~kind_b, ca:] = 0
    feats[~kind_b, 0] = 1
    feats[kind_b, ca] = 1
    filt = np.zeros(ks + (cin, cout), np.float32)
    filt[..., :ca, :na] = rng.uniform(-0.5, 0.5, ks + (ca, na))
    filt[..., ca:, na:na + nb] = rng.uniform(-0.5, 0.5, ks + (cb, nb))
    dk = np.zeros((cin, cout), np.float32)
    dk[:, na + nb:] = rng.uniform(-0.5, 0.5, (cin, n_dense))
    bias = rng.standard_normal(cout).astype(np.float32)
    extent = np.float32(0.3)
    radius = np.float32(0.5) * extent
    idx, splits, d2 = o64.fixed_radius_search(pts, pts, radius, ignore_query_point=True)
    imp = o64.window("poly6", d2.astype(np.float64) / np.float64(radius) ** 2)
    ref = o64.continuous_conv(filt, pts, extent, (0, 0, 0), pts, feats * np.float32(0.25), None, idx, imp, splits,
                              align_corners=True, coordinate_mapping="ball_to_cube_volume_preserving", normalize=False,
                              interpolation="linear")
    ref = ref + feats.astype(np.float64) @ dk + bias
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    w_ext = torch.cat([t(filt).reshape(-1, cout), t(dk)], dim=0)
    recs = None
    if with_records:
        recs = ops.prepare_pair_records(ks, t(pts), float(extent), None, t(pts), None, t(idx), None, t(splits), align_corners=True,
                                        coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", window="poly6")
    got = {}
    for opt in (3, 3 | 4096):
        prev = ops.set_kernel_options(opt)
        try:
            launches = ops.launch_count()
            got[opt] = ops.continuous_conv(
                w_ext, t(pts), float(extent), None, t(pts), t(feats), None, t(idx), None, t(splits), align_corners=True,
                coordinate_mapping="ball_to_cube_volume_preserving", normalize=False, interpolation="linear", window="poly6",
                feat_scale=0.25, bias=t(bias), dense_inp=t(feats), dense_cin=cin, kernel_size=ks, pair_records=recs,
                block_diagonal=(ca, na, nb)).cpu().numpy()
            assert ops.launch_count() == launches + 1
        finally:
            ops.set_kernel_options(prev)
        feat_close(got[opt], ref)
    feat_close(got[3], got[3 | 4096], 0.5)
    # a promise that does not fit the layer is refused
    with pytest.raises(Exception):
        ops.continuous_conv(w_ext, t(pts), float(extent), None, t(pts), t(feats), None, t(idx), None, t(splits), align_corners=True,
                            coordinate_mapping="ball_to_cube_volume_preserving", window="poly6", dense_inp=t(feats), dense_cin=cin,
                            kernel_size=ks, block_diagonal=(cin, na, nb))


def test_ascc_fused_matches_reference_form_and_conserves(cuda, kernel_options):
    """Antisymmetric layer: fused (f_j + f_i) kernel vs the reference's two-pass form (utils/convolutions.py:433-458)
    and momentum conservation sum_i out_i = 0."""
    from dmcf_b200 import ops
    rng = np.random.default_rng(13)
    n, cin, cout = 1500, 32, 3
    pts = rng.random((n, 3)).astype(np.float32) * 0.6
    feats = rng.standard_normal((n, cin)).astype(np.float32)
    half = rng.uniform(-0.5, 0.5, (6, 3, 6, cin, cout)).astype(np.float32)
    extent = np.float32(0.2)
    ref = o64.cconv_layer(np.maximum(feats, 0), pts, pts, extent, half, None, align_corners=True,
                          coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", normalize=False,
                          ignore_query_points=True, window_name="peak", symmetric=True, sym_axis=1)
    full = o64.symmetric_kernel(half, 1).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    nns = ops.fixed_radius_search(t(pts), t(pts), float(np.float32(0.5) * extent), ignore_query_point=True)
    got = ops.continuous_conv(t(full), t(pts), float(extent), None, t(pts), t(feats), None, nns.neighbors_index, None,
                              nns.neighbors_row_splits, align_corners=True,
                              coordinate_mapping="ball_to_cube_volume_preserving", normalize=False, interpolation="linear",
                              window="peak", relu_input=True, ascc=True).cpu().numpy().astype(np.float64)
    feat_close(got, ref)
    total = np.abs(got).sum(axis=0)
    assert np.all(np.abs(got.sum(axis=0)) <= 1e-5 * total + 1e-5), (got.sum(axis=0), total)
    # the same layer with the antisymmetry promise (folded half-patch kernel k_cconv_apatch), also through pair records
    recs = ops.prepare_pair_records((6, 6, 6), t(pts), float(extent), None, t(pts), None, nns.neighbors_index, None,
                                    nns.neighbors_row_splits, align_corners=True,
                                    coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", window="peak")
    for rec in (None, recs):
        got2 = ops.continuous_conv(t(full), t(pts), float(extent), None, t(pts), t(feats), None, nns.neighbors_index, None,
                                   nns.neighbors_row_splits, align_corners=True,
                                   coordinate_mapping="ball_to_cube_volume_preserving", normalize=False,
                                   interpolation="linear", window="peak", relu_input=True, ascc=True,
                                   antisymmetric_filter=True, pair_records=rec).cpu().numpy().astype(np.float64)
        feat_close(got2, ref)
        feat_close(got2, got, 0.5)
        assert np.all(np.abs(got2.sum(axis=0)) <= 1e-5 * total + 1e-5), (got2.sum(axis=0), total)


def test_ascc_2d_half_patch(cuda, kernel_options):
    """2-D antisymmetric layer (WBC-SPH: stored [1,4,8,32,2], effective 1x8x8, sym_axis 1), long rows, fused Dense, bias,
    residual: folded half-patch kernel vs the oracle's two-pass form."""
    from dmcf_b200 import ops
    rng = np.random.default_rng(29)
    n, cin, cout = 1200, 32, 2
    pts = rng.random((n, 3)).astype(np.float32) * 0.5
    pts[:, 2] = 0
    feats = rng.standard_normal((n, cin)).astype(np.float32)
    half = rng.uniform(-0.5, 0.5, (1, 4, 8, cin, cout)).astype(np.float32)
    extent = np.float32(0.12)
    ref = o64.cconv_layer(np.maximum(feats, 0), pts, pts, extent, half, None, align_corners=True,
                          coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", normalize=False,
                          ignore_query_points=True, window_name="peak", symmetric=True, sym_axis=1)
    dk = rng.uniform(-0.3, 0.3, (cin, cout)).astype(np.float32)
    bias = rng.standard_normal(cout).astype(np.float32)
    resid = rng.standard_normal((n, cout)).astype(np.float32)
    ref = ref + np.maximum(feats, 0).astype(np.float64) @ dk + bias + resid
    full = o64.symmetric_kernel(half, 1).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    nns = ops.fixed_radius_search(t(pts), t(pts), float(np.float32(0.5) * extent), ignore_query_point=True)
    assert int(np.diff(nns.neighbors_row_splits.cpu().numpy()).max()) > 32  # several chunks per row
    w_ext = torch.cat([t(full).reshape(-1, cout), t(dk)], dim=0)
    got = ops.continuous_conv(w_ext, t(pts), float(extent), None, t(pts), t(feats), None, nns.neighbors_index, None,
                              nns.neighbors_row_splits, align_corners=True,
                              coordinate_mapping="ball_to_cube_volume_preserving", normalize=False, interpolation="linear",
                              window="peak", relu_input=True, ascc=True, antisymmetric_filter=True, bias=t(bias),
                              dense_inp=t(feats), dense_cin=cin, residual=t(resid), kernel_size=(1, 8, 8)).cpu().numpy()
    feat_close(got, ref)


def test_dense_and_elementwise(cuda):
    from dmcf_b200 import ops
    rng = np.random.default_rng(17)
    for n, cin, cout in [(1000, 7, 8), (513, 32, 32), (77, 96, 64), (1, 4, 8), (0, 4, 8)]:
        x = rng.standard_normal((n, cin)).astype(np.float32)
        w = rng.standard_normal((cin, cout)).astype(np.float32)
        b = rng.standard_normal(cout).astype(np.float32)
        t = lambda a: torch.from_numpy(a).to(cuda)
        got = ops.dense(t(x), t(w), t(b), relu_input=True).cpu().numpy()
        feat_close(got, o64.dense(np.maximum(x, 0), w, b))
    n = 1234
    pos = rng.random((n, 3)).astype(np.float32); vel = rng.standard_normal((n, 3)).astype(np.float32)
    acc = rng.standard_normal((n, 3)).astype(np.float32)
    dt = np.float32(0.02)
    t = lambda a: torch.from_numpy(a).to(cuda)
    p2, v2 = ops.integrate(t(pos), t(vel), t(acc), (0, -9.81, 0), float(dt))
    v_ref = vel + dt * acc
    p_ref = pos + dt * v_ref
    assert np.allclose(v2.cpu().numpy(), v_ref, rtol=1e-6, atol=1e-7) and np.allclose(p2.cpu().numpy(), p_ref, rtol=1e-6, atol=1e-7)
    p2g, v2g = ops.integrate(t(pos), t(vel), None, (0, -9.81, 0), float(dt))
    assert np.allclose(v2g.cpu().numpy(), vel + dt * np.array([0, -9.81, 0], np.float32), rtol=1e-6, atol=1e-7)
    for c in (1, 2, 3):
        net = rng.standard_normal((n + 50, c)).astype(np.float32)
        pn, vn = ops.correct(t(pos), p2, t(net), (0.5, 0.25, 0.125), float(dt))
        e = np.repeat(net, 3, axis=1) if c == 1 else (np.concatenate([net, net[:, :1]], 1) if c == 2 else net)
        pr = p2.cpu().numpy() + np.array([0.5, 0.25, 0.125], np.float32) * e[:n]
        assert np.allclose(pn.cpu().numpy(), pr, rtol=1e-6, atol=1e-7)
        assert np.allclose(vn.cpu().numpy(), (pr - pos) / dt, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n,cin,cout", [(10_000, 32, 32), (4096, 4, 8), (70_001, 96, 64), (5000, 24, 3), (33_333, 7, 20),
                                        (200_000, 32, 32)])
def test_dense_tensor_core_kernel(cuda, n, cin, cout):
    """k_dense_umma (tcgen05.mma kind::tf32 as 3xTF32, accumulator in tensor memory) against the float64 oracle and against the
    SIMT kernel k_dense (option bit 13), with relu, strided input / output rows and a ragged last tile."""
    from dmcf_b200 import ops
    rng = np.random.default_rng(n + cin)
    x_wide = rng.standard_normal((n, cin + 8)).astype(np.float32)
    x = x_wide[:, 4:4 + cin]
    w = rng.standard_normal((cin, cout)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    ref = o64.dense(np.maximum(x, 0), w, b)
    xw = t(x_wide)
    got = {}
    for opt in (3, 3 | 8192):
        prev = ops.set_kernel_options(opt)
        try:
            out_wide = torch.full((n, cout + 4), -7.0, device=cuda)
            ops.dense(xw[:, 4:4 + cin], t(w), t(b), relu_input=True, out=out_wide[:, 4:])
            got[opt] = out_wide[:, 4:].cpu().numpy()
            assert float((out_wide[:, :4] + 7.0).abs().sum()) == 0
        finally:
            ops.set_kernel_options(prev)
        feat_close(got[opt], ref)
    feat_close(got[3], got[3 | 8192], 0.5)
    # contiguous rows, no relu, no bias
    plain = ops.dense(t(x), t(w), None).cpu().numpy()
    feat_close(plain, o64.dense(x, w, np.zeros(cout, np.float32)))


def test_no_cpu_fallback(cuda):
    from dmcf_b200 import ops
    from dmcf_b200._lib import DmcfError
    with pytest.raises(DmcfError):
        ops.fixed_radius_search(torch.zeros((4, 3)), torch.zeros((4, 3)), 0.1)


@pytest.mark.parametrize("ks,cin,cout,extent", [((4, 4, 4), 4, 16, 0.45), ((4, 4, 4), 8, 32, 0.45), ((4, 4, 4), 16, 8, 0.45),
                                                ((1, 8, 8), 5, 12, 0.3), ((1, 8, 1), 2, 4, 0.2), ((4, 4, 4), 1, 4, 0.4)])
def test_multipair_phase1_cross_sets_long_rows(cuda, ks, cin, cout, extent):
    """Multi-pair phase 1 (cin <= 16) on the shape of the cross-scale convs: distinct in / out sets, rows of 100+ pairs
    (several chunks, the chunk pipeline), rows without any neighbour, an out-point count that is not a multiple of the
    24-point tile, with and without pair records, the antisymmetric centre term on a 4-channel output; against the oracle
    and against the single-pair walk."""
    from dmcf_b200 import ops
    rng = np.random.default_rng(cin * 100 + cout)
    n_in, n_out = 2600, 1013
    pts = rng.random((n_in, 3)).astype(np.float32)
    outp = (rng.random((n_out, 3)) * 1.3 - 0.15).astype(np.float32)  # some out points have no neighbour at all
    if ks[0] == 1:
        pts[:, 2] = 0; outp[:, 2] = 0
    if ks[2] == 1:
        pts[:, 0] = 0; outp[:, 0] = 0
    outp[:3] = np.where(np.array(ks[::-1]) > 1, 5.0, 0.0).astype(np.float32)  # far away: rows without neighbours
    feats = rng.standard_normal((n_in, cin)).astype(np.float32)
    filt = rng.uniform(-0.5, 0.5, ks + (cin, cout)).astype(np.float32)
    ext = np.float32(extent)
    radius = np.float32(0.5) * ext
    idx, splits, d2 = o64.fixed_radius_search(pts, outp, radius)
    counts = np.diff(splits)
    assert counts.max() > 64 and counts.min() == 0
    imp = o64.window("poly6", d2.astype(np.float64) / (np.float64(radius) ** 2))
    ref = o64.continuous_conv(filt, outp, ext, (0, 0, 0), pts, np.maximum(feats, 0) * 0.7, None, idx, imp, splits,
                              align_corners=True, coordinate_mapping="ball_to_cube_volume_preserving", normalize=False,
                              interpolation="linear")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    kw = dict(align_corners=True, coordinate_mapping="ball_to_cube_volume_preserving", normalize=False, interpolation="linear",
              window="poly6", relu_input=True, feat_scale=0.7)
    recs = ops.prepare_pair_records(ks, t(outp), float(ext), None, t(pts), None, t(idx), None, t(splits), align_corners=True,
                                    coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", window="poly6")
    outs = {}
    for opt in (3, 35):
        prev = ops.set_kernel_options(opt)
        try:
            for rec in (None, recs):
                got = ops.continuous_conv(t(filt), t(outp), float(ext), None, t(pts), t(feats), None, t(idx), None, t(splits),
                                          pair_records=rec, **kw).cpu().numpy()
                feat_close(got, ref)
                outs[(opt, rec is None)] = got
        finally:
            ops.set_kernel_options(prev)
    feat_close(outs[(3, True)], outs[(35, True)], 0.5)
    feat_close(outs[(3, False)], outs[(35, False)], 0.5)
    assert np.all(outs[(3, True)][counts == 0] == 0)
    # normaliser (sum of the window values) through the multi-pair path
    refn = o64.continuous_conv(filt, outp, ext, (0, 0, 0), pts, feats, None, idx, imp, splits, align_corners=True,
                               coordinate_mapping="ball_to_cube_volume_preserving", normalize=True, interpolation="linear")
    gotn = ops.continuous_conv(t(filt), t(outp), float(ext), None, t(pts), t(feats), None, t(idx), None, t(splits),
                               align_corners=True, coordinate_mapping="ball_to_cube_volume_preserving", normalize=True,
                               interpolation="linear", window="poly6").cpu().numpy()
    feat_close(gotn, refn)


@pytest.mark.parametrize("cin", [4, 8, 16])
def test_multipair_phase1_antisymmetric_centre_term(cuda, cin):
    """ascc on a 4-channel output that takes the k_cconv_lean path (apatch / direct switched off): out_i = sum_j W(r_ij)
    (f_j + f_i) with the centre term added per pair slot; against the single-pair walk and momentum conservation."""
    from dmcf_b200 import ops
    rng = np.random.default_rng(40 + cin)
    n, cout = 1100, 4
    pts = rng.random((n, 3)).astype(np.float32) * 0.6
    feats = rng.standard_normal((n, cin)).astype(np.float32)
    half = rng.uniform(-0.5, 0.5, (4, 2, 4, cin, cout)).astype(np.float32)
    full = o64.symmetric_kernel(half, 1).astype(np.float32)
    extent = np.float32(0.25)
    ref = o64.cconv_layer(np.maximum(feats, 0), pts, pts, extent, half, None, align_corners=True,
                          coordinate_mapping="ball_to_cube_volume_preserving", interpolation="linear", normalize=False,
                          ignore_query_points=True, window_name="peak", symmetric=True, sym_axis=1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    nns = ops.fixed_radius_search(t(pts), t(pts), float(np.float32(0.5) * extent), ignore_query_point=True)
    outs = []
    for opt in (1, 1 | 32):  # bit 1 off: no direct / apatch kernel, the lean kernel takes cout = 4
        prev = ops.set_kernel_options(opt)
        try:
            got = ops.continuous_conv(t(full), t(pts), float(extent), None, t(pts), t(feats), None, nns.neighbors_index, None,
                                      nns.neighbors_row_splits, align_corners=True,
                                      coordinate_mapping="ball_to_cube_volume_preserving", normalize=False,
                                      interpolation="linear", window="peak", relu_input=True, ascc=True).cpu().numpy()
        finally:
            ops.set_kernel_options(prev)
        feat_close(got, ref)
        total = np.abs(got).sum(axis=0)
        assert np.all(np.abs(got.astype(np.float64).sum(axis=0)) <= 1e-5 * total + 1e-5)
        outs.append(got)
    feat_close(outs[0], outs[1], 0.5)
