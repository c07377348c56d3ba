"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys (the CPU port timed on
the host cores), the GPU arm refuses to run without a CUDA device (no CPU fallback), and the helper that turns the live
CUDA-event records of the HBM-bound ops into GB/s does its arithmetic."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must set its thread count itself
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-n-side", "12"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stderr[-500:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["vs_baseline"] is None and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and "sample" in line["cpu_baseline"]
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert line["scaling"] == "strong" and line["config"]["same_config_as_gpu_arm"] is False  # --cpu-n-side given: a sub-volume
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert res.returncode != 0 and "no CPU fallback" in (res.stderr + res.stdout)


def test_hbm_op_breakdown_arithmetic():
    sys.path.insert(0, ROOT)
    import bench

    class Ev:
        def __init__(self, t):
            self.t = t

        def elapsed_time(self, other):
            return other.t - self.t

    recs = [dict(kind="frs_fill", n_points=1000, n_queries=500, pairs=20000, distances=False, start=Ev(0.0), end=Ev(0.5)),
            dict(kind="frs_fill", n_points=1000, n_queries=500, pairs=20000, distances=True, start=Ev(1.0), end=Ev(1.5)),
            dict(kind="pair_records", n_inp=1000, n_out=500, pairs=20000, start=Ev(0.0), end=Ev(2.0))]
    out = {r["kernel"]: r for r in bench.hbm_op_breakdown(recs, 10.0, 1000.0)}
    fill = out["k_frs<fill>"]
    bytes_fill = 2 * (12 * 1500 + 8 * 501) + 4 * 20000 + 8 * 20000
    assert fill["launches"] == 2 and abs(fill["avg_ms"] - 0.5) < 1e-9 and abs(fill["share_of_step"] - 0.1) < 1e-9
    assert abs(fill["GBps"] - round(bytes_fill / 1e9 / 1e-3, 1)) < 0.11
    prep = out["k_cconv_prepare"]
    assert abs(prep["GBps"] - round((40 * 20000 + 12 * 1500 + 8 * 501) / 1e9 / 2e-3, 1)) < 0.11
    assert abs(prep["frac_of_hbm_peak"] - prep["GBps"] / 1000.0) < 1e-3
