"""Timings of the point-set ops on one GPU: CUDA events, 3 warm-up + 5 timed calls each.
Writes gpurun_out/pointset_bench.json.  `python scripts/bench_pointset.py`"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dmcf_b200 import pointops  # noqa: E402


def timed(fn, warm=2, reps=4):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(0)
    out = []
    for n, m in [(10408, 5204), (108856, 54428), (108856, 13607)]:  # WBC-SPH / Liquid3d sized sets, strides 2 and 8
        pts = torch.rand((1, n, 3), generator=gen).to(dev)
        for cl in (1, 2, 4, 8):
            best, avg = timed(lambda: pointops.farthest_point_sample(m, pts, cluster_size=cl), 1, 3)
            out.append({"op": "farthest_point_sample", "n": n, "m": m, "cluster": cl, "ms_best": best, "ms_avg": avg,
                        "us_per_sample": 1e3 * best / m})
            print(out[-1], flush=True)
    for n in (10000, 30000, 100000):
        a = (torch.rand((1, n, 3), generator=gen) * 0.5).to(dev)
        b = (a + 0.004 * torch.randn((1, n, 3), generator=gen).to(dev)).contiguous()
        best, avg = timed(lambda: pointops.emd_cost(a, b), 1, 3)
        pairs = 10 * 3 * n * n
        out.append({"op": "emd_cost (fused, no match matrix)", "n": n, "m": n, "ms_best": best, "ms_avg": avg,
                    "pair_evals_per_s": pairs / (best * 1e-3)})
        print(out[-1], flush=True)
        if n * n * 4 <= 8 << 30:
            best, avg = timed(lambda: pointops.match_cost(a, b, pointops.approx_match(a, b)), 1, 3)
            out.append({"op": "approx_match + match_cost", "n": n, "m": n, "ms_best": best, "ms_avg": avg,
                        "match_bytes": n * n * 4})
            print(out[-1], flush=True)
        best, avg = timed(lambda: pointops.nn_distance(a, b), 1, 3)
        out.append({"op": "nn_distance (both directions)", "n": n, "m": n, "ms_best": best, "ms_avg": avg,
                    "pair_evals_per_s": 2 * n * n / (best * 1e-3)})
        print(out[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/pointset_bench.json", "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
