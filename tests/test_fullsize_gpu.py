"""BASELINE's full size (config 4: 100^3 = 1 M fluid particles + 62 k wall particles, 32 M neighbour pairs) through properties
that do not need the oracle to run at that size: symmetry of the neighbour relation and exact agreement with a brute-force
search on sampled rows, linearity and momentum conservation of the conv layers, determinism, and the graph step against the
eager step.  The small-size parity tests (tests/test_ops_gpu.py, test_models_gpu.py) hold the kernels to the oracle; these hold
the same kernels to the mathematics at the size the bench times."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c4(cuda):
    from dmcf_b200 import config, scenes
    scene = scenes.lattice_scene((100, 100, 100), dx=0.05, seed=0)
    model = config.build_model(scenes.c4_model_cfg())
    model.init_weights(seed=0, device=cuda, scale=0.1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(cuda)
    sample = [t(scene["pos"]), t(scene["vel"]), None, None, t(scene["box"]), t(scene["box_normals"])]
    all_pos = torch.cat([sample[0], sample[4]], dim=0).contiguous()
    return scene, model, sample, all_pos


def test_search_is_symmetric_and_matches_brute_force_rows(cuda, c4):
    from dmcf_b200 import ops
    _, _, _, pts = c4
    n = pts.shape[0]
    radius = 0.1
    nns = ops.fixed_radius_search(pts, pts, radius)
    rs = nns.neighbors_row_splits
    idx = nns.neighbors_index.to(torch.int64)
    p = int(rs[-1])
    assert p == idx.shape[0] and 30_000_000 < p < 36_000_000
    rows = torch.repeat_interleave(torch.arange(n, device=cuda), rs[1:] - rs[:-1])
    # j in N(i) <=> i in N(j): every moment of the row index over the pairs equals that of the neighbour index
    assert int(rows.sum()) == int(idx.sum())
    assert int((rows * rows % 1_000_003).sum()) == int((idx * idx % 1_000_003).sum())
    # every point is its own neighbour exactly once
    assert int((rows == idx).sum()) == n
    # sampled rows against a brute-force search with the oracle's operation order, (dx*dx + dy*dy) + dz*dz in float32
    g = torch.Generator(device="cpu").manual_seed(3)
    r2 = torch.tensor(radius, dtype=torch.float32, device=cuda) ** 2
    for q in torch.randint(0, n, (64,), generator=g).tolist():
        d = pts - pts[q]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        ref = torch.nonzero(d2 <= r2).flatten()
        got = torch.sort(idx[int(rs[q]):int(rs[q + 1])]).values
        assert torch.equal(got, ref), q


def test_conv_layers_are_linear_deterministic_and_conserve_momentum(cuda, c4):
    from dmcf_b200 import ops
    _, _, _, pts = c4
    n = pts.shape[0]
    g = torch.Generator(device="cpu").manual_seed(5)
    f1 = torch.randn((n, 32), generator=g).to(cuda)
    f2 = torch.randn((n, 32), generator=g).to(cuda)
    w = (torch.rand((4, 4, 4, 32, 32), generator=g) - 0.5).to(cuda)
    nns = ops.fixed_radius_search(pts, pts, 0.1)
    kw = dict(align_corners=True, coordinate_mapping="ball_to_cube_volume_preserving", normalize=False, interpolation="linear",
              window="poly6")
    conv = lambda f, filt, **extra: ops.continuous_conv(filt, pts, 0.2, None, pts, f, None, nns.neighbors_index, None,
                                                        nns.neighbors_row_splits, **kw, **extra)
    a = conv(f1, w)
    assert torch.equal(a, conv(f1, w)), "the wide-layer kernel is not deterministic"
    lin = conv(1.5 * f1 + f2, w)
    ref = 1.5 * a + conv(f2, w)
    scale = float(ref.abs().max())
    assert float((lin - ref).abs().max()) <= 2e-5 * scale, "conv is not linear in the features"
    # antisymmetric 6x6x6 layer with the fused centre term: sum_i out_i = 0 (momentum conservation, utils/convolutions.py:433-458)
    half = (torch.rand((6, 3, 6, 32, 3), generator=g) - 0.5).to(cuda)
    full = torch.cat([-torch.flip(half, dims=(0, 1, 2)), half], dim=1).contiguous()  # utils/convolutions.py:410-412
    out = conv(f1, full, ascc=True, skip_self=True, relu_input=True, antisymmetric_filter=True)
    total = out.to(torch.float64).sum(dim=0).abs().max()
    mass = out.to(torch.float64).abs().sum(dim=0).max()
    assert float(total) <= 1e-5 * float(mass), (float(total), float(mass))
    assert torch.equal(out, conv(f1, full, ascc=True, skip_self=True, relu_input=True, antisymmetric_filter=True))


def test_graph_step_matches_the_eager_step_at_full_size(cuda, c4):
    from dmcf_b200.simulator import Simulator
    _, model, sample, _ = c4
    eager = Simulator(model, device="cuda", step_mode="eager")
    sim = Simulator(model, device="cuda", step_mode="graph")
    with torch.no_grad():
        ref = eager.step(sample)
        outs = [sim.step(sample) for _ in range(4)]
    assert sim.stats["graph_replays"] == 2 and sim.stats["replans"] == 0, sim.stats
    from dmcf_b200 import scenes
    n = sample[0].shape[0]
    dt = float(scenes.c4_model_cfg()["timestep"])
    scale = max(1.0, float(ref[0].abs().max()))
    for o in outs:
        assert o[0].shape[0] == n and bool(torch.isfinite(o[0]).all()) and bool(torch.isfinite(o[1]).all())
        # replays use padded cell grids (another summation order inside a neighbour row): rounding only
        assert float((o[0] - ref[0]).abs().max()) <= 2e-6 * scale
        assert float((o[1] - ref[1]).abs().max()) <= 2e-6 * scale / dt * 2
    assert torch.equal(outs[-1][0], outs[-2][0]), "graph replays differ"
    # one step moves a particle by a fraction of the domain (a diverging kernel would show here)
    assert float((outs[-1][0] - sample[0]).abs().max()) < 0.5
