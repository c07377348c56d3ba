"""Host-side validation metrics (dmcf_b200/metrics.py vs utils/evaluation_helper.py:14-90)."""
import numpy as np

from dmcf_b200 import metrics


def _compare_dist_loop(x, y, bin_size=25):
    """the reference's per-sample loop (utils/evaluation_helper.py:43-72), restated for the test"""
    from scipy.stats import entropy
    cnt, dim = x.shape[0], x.shape[-1]
    b = int((cnt // bin_size) ** (1 / dim))
    both = np.concatenate((x, y), axis=0)
    mn, mx = np.percentile(both, 5, axis=0), np.percentile(both, 95, axis=0)
    w = (mx - mn + 1e-6) / b
    hx, hy = np.zeros((b + 1,) * dim) + 1e-5, np.zeros((b + 1,) * dim) + 1e-5
    idx = lambda v: tuple(np.clip(((v - mn) / w).astype("int32"), 0, b))
    for v in x:
        hx[idx(v)] += 1
    for v in y:
        hy[idx(v)] += 1
    return entropy(hx.reshape(-1), hy.reshape(-1))


def test_metrics_match_reference_definitions():
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((500, 3)), rng.standard_normal((500, 3)) * 1.2 + 0.1
    assert np.allclose(metrics.distance(a, b), np.sqrt(((a - b) ** 2).sum(-1)))
    ch = metrics.chamfer_distance(a, b)
    brute = np.sqrt(((b[:, None] - a[None]) ** 2).sum(-1)).min(1)
    assert np.allclose(ch, brute)
    assert np.isclose(metrics.compare_dist(a, b), _compare_dist_loop(a, b))
    assert metrics.compare_dist(a, a) < 1e-12
    s = metrics.compute_stats(ch)
    assert s["num_particles"] == 500 and np.isclose(s["mse"], np.mean(ch ** 2))
    m = metrics.merge_dicts([{"x": 1.0}, {"x": 3.0}], lambda p, q: p + q / 2)
    assert np.isclose(m["x"], 2.0)
