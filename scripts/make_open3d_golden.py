"""Writes tests/golden/open3d_conv_cases.npz: the outputs of the REAL Open3D ops for the seeded continuous_conv parity cases
(tests/conv_cases.py).  This is the exit from "parity unpinned" at the Open3D boundary: the arithmetic
of `fixed_radius_search` / `continuous_conv` lives in the open3d wheel (reference pin: open3d==0.15.2, requirements.txt:2),
which cannot be installed in the build container (Python 3.12, no network).  On ANY machine where `open3d.ml.torch` or
`open3d.ml.tf` imports (CPU is enough):

    python scripts/make_open3d_golden.py            # writes tests/golden/open3d_conv_cases.npz
    python -m pytest tests/test_oracle_cpu.py -k open3d   # oracle O64 / O32 against it (and tests/test_ops_gpu.py on a GPU)

The file holds, per case: neighbors_index / neighbors_row_splits / neighbors_distance of
ml3d.layers.FixedRadiusSearch(metric='L2', ignore_query_point, return_distances=True)(points, queries, radius) -- the call of
utils/convolutions.py:354-358 -- and the output of ml3d.ops.continuous_conv with the kwargs assembled at
utils/convolutions.py:414-429 (window importances computed like :359-379)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_backend():
    try:
        import torch
        import open3d.ml.torch as ml3d
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        back = lambda t: t.detach().cpu().numpy()
        return "torch", ml3d, to, back
    except Exception as e_torch:  # noqa: BLE001
        try:
            import tensorflow as tf
            import open3d.ml.tf as ml3d
            to = lambda a: tf.convert_to_tensor(np.ascontiguousarray(a))
            back = lambda t: t.numpy()
            return "tf", ml3d, to, back
        except Exception as e_tf:  # noqa: BLE001
            raise SystemExit(f"neither open3d.ml.torch ({e_torch}) nor open3d.ml.tf ({e_tf}) is importable here; "
                             "run this script where the open3d wheel is installed")


def main():
    from conv_cases import CONV_CASES, conv_case_inputs
    from oracle import o64  # only for the window importances, which are DMCF code (utils/tools/losses.py:8-44), not Open3D's
    name, ml3d, to, back = load_backend()
    import open3d
    out = {"open3d_version": np.asarray(open3d.__version__), "backend": np.asarray(name)}
    for case in CONV_CASES:
        ks, cin, cout, mapping, interp, align, normalize, window, ignore_q, pts, outp, feats, filt, extent, radius = conv_case_inputs(case)
        frs = ml3d.layers.FixedRadiusSearch(metric="L2", ignore_query_point=bool(ignore_q), return_distances=True)
        nns = frs(to(pts), to(outp), float(radius))
        idx, splits, d2 = back(nns.neighbors_index), back(nns.neighbors_row_splits), back(nns.neighbors_distance)
        imp = np.zeros(0, np.float32)
        if window is not None:
            imp = o64.window(window, d2.astype(np.float64) / (np.float64(radius) ** 2)).astype(np.float32)
        res = ml3d.ops.continuous_conv(
            filters=to(filt), out_positions=to(outp), extents=to(np.asarray([extent], np.float32)),
            offset=to(np.zeros(3, np.float32)), inp_positions=to(pts), inp_features=to(feats),
            inp_importance=to(np.zeros(0, np.float32)), neighbors_index=nns.neighbors_index,
            neighbors_importance=to(imp), neighbors_row_splits=nns.neighbors_row_splits, align_corners=bool(align),
            coordinate_mapping=mapping, normalize=bool(normalize), interpolation=interp, max_temp_mem_MB=64)
        c = case[0]
        out[c + "/index"], out[c + "/row_splits"], out[c + "/distance"], out[c + "/out"] = idx, splits, d2, back(res)
        print(c, "pairs", len(idx), "out", back(res).shape)
    path = os.path.join(ROOT, "tests", "golden", "open3d_conv_cases.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
