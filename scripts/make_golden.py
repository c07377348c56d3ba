"""Generates the committed fixtures under tests/golden/ from the reference checkout (run HERE, where
/root/reference exists; the GPU box only sees the generated files).

  ckpt_<name>.npz       trained weights of the shipped checkpoints, read with dmcf_b200.checkpoint (TF-free)
  column_seed44.npz     1-D SPH column from the reference's own generator datasets/column_gen.py (imported, not copied),
                        with the generator's brute-force neighbour counter SPH1D.cnt_nn (:36-43) as the known answer
  canyon_crop.npz       frame 0 of datasets/canyon_data/canyon.msgpack.zst cropped around the inflow block, plus the particle
                        positions of all 13 shipped ground-truth frames (the block falls and hits the canyon floor)
  pointset_ref.npz      outputs of the reference's own CPU functions for the in-repo point-set ops (approxmatch_cpu,
                        approxmatch_cpu_dyn, matchcost_cpu, matchcostgrad_cpu of utils/tools/tf_approxmatch.cpp, nnsearch of
                        utils/tools/nn_distance.cpp), run through oracle/_ref/libdmcf_refops.so (built by oracle/Makefile from
                        the reference sources where they lie) on seeded point clouds
"""
import os
import sys

import numpy as np

REF = os.environ.get("DMCF_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def checkpoints():
    from dmcf_b200.checkpoint import load_checkpoint, model_weights
    for name in ("Liquid3d", "WBC-SPH"):
        w = model_weights(load_checkpoint(os.path.join(REF, "checkpoints", name, "ckpt")))
        np.savez_compressed(os.path.join(OUT, f"ckpt_{name}.npz"), **{k.replace("/", "|"): v for k, v in w.items()})
        print(name, len(w), "tensors", sum(v.size for v in w.values()), "params")


def column():
    import importlib.util  # load the module file directly: datasets/__init__.py pulls in tensorpack
    spec = importlib.util.spec_from_file_location("ref_column_gen", os.path.join(REF, "datasets", "column_gen.py"))
    column_gen = importlib.util.module_from_spec(spec)  # the reference's own generator (pure NumPy)
    spec.loader.exec_module(column_gen)
    np.random.seed(44)  # datasets/dataset_reader_physics.py:148-151 seeds the generator like this
    data = column_gen.gen_data(data_cnt=1, timesteps=12, pts_cnt=[40], res=100, dt=0.0025, gravity=-10.0, width=1)
    frames = data[0]
    pos = np.stack([f["pos"] for f in frames]).astype(np.float32)
    vel = np.stack([f["vel"] for f in frames]).astype(np.float32)
    box = np.stack([f["box"] for f in frames]).astype(np.float32)
    # known answer: the generator's own neighbour counter on the solver state (1-D positions in solver units)
    solver = column_gen.SPH1D(radius=0.25, mass=1.0, stiffness=20.0, visc=0.1, gravity=-10.0 * 100)
    np.random.seed(7)
    solver.setup(60, 2, rnd=0.15)
    solver_x = solver.particles[:, 0].astype(np.float32).copy()
    cnt = solver.cnt_nn().astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "column_seed44.npz"), pos=pos, vel=vel, box=box,
                        box_normals=np.stack([f["box_normals"] for f in frames]).astype(np.float32),
                        grav=np.asarray(frames[0]["grav"], np.float32), solver_x=solver_x, solver_h=np.float32(solver.h),
                        solver_cnt_nn=cnt)
    print("column", pos.shape, box.shape, "cnt_nn", cnt[:8])


def canyon():
    import msgpack
    import pyarrow
    raw = open(os.path.join(REF, "datasets", "canyon_data", "canyon.msgpack.zst"), "rb").read()
    buf = pyarrow.CompressedInputStream(pyarrow.BufferReader(raw), "zstd").read()

    def hook(d):
        d = {(k.decode() if isinstance(k, bytes) else k): v for k, v in d.items()}
        if "nd" in d and "type" in d and "data" in d:
            t = d["type"].decode() if isinstance(d["type"], bytes) else d["type"]
            a = np.frombuffer(d["data"], dtype=np.dtype(t))
            return a.reshape(d["shape"]) if d["nd"] else a[0]
        return d

    frames = msgpack.unpackb(buf, raw=True, object_hook=hook, strict_map_key=False)
    f0 = frames[0]
    pos, vel = f0["pos"], f0["vel"]
    lo, hi = pos.min(0) - 0.7, pos.max(0) + 0.7
    m = np.all((f0["box"] >= lo) & (f0["box"] <= hi), axis=1)
    np.savez_compressed(os.path.join(OUT, "canyon_crop.npz"), pos=pos.astype(np.float32), vel=vel.astype(np.float32),
                        box=f0["box"][m].astype(np.float32), box_normals=f0["box_normals"][m].astype(np.float32),
                        gt_pos_1=frames[1]["pos"].astype(np.float32),
                        gt_pos=np.stack([f["pos"] for f in frames]).astype(np.float32))  # the 13 SPH ground-truth frames
    print("canyon", pos.shape, int(m.sum()), "boundary points kept of", len(m))


def pointset():
    from oracle import pointset as ps
    assert ps.ref_available(), "build oracle/_ref first: make -C oracle"
    out = {}
    cases = [("a", 96, 96, 1.0, 3), ("b", 150, 75, 0.3, 3), ("c", 60, 140, 0.05, 3), ("d", 64, 64, 0.02, 2)]
    for name, n, m, scale, dim in cases:
        rng = np.random.default_rng(sum(map(ord, name)) + n)
        x1 = (rng.random((n, 3)) * scale).astype(np.float32)
        x2 = (rng.random((m, 3)) * scale).astype(np.float32)
        if dim == 2:
            x1[:, 2] = 0
            x2[:, 2] = 0
        match = ps.ref_approx_match(x1[None], x2[None])
        out[f"{name}_xyz1"], out[f"{name}_xyz2"], out[f"{name}_match"] = x1, x2, match[0]
        out[f"{name}_cost"] = ps.ref_match_cost(x1[None], x2[None], match)[0]
        g1, g2 = ps.ref_match_cost_grad(x1[None], x2[None], match)
        out[f"{name}_grad1"], out[f"{name}_grad2"] = g1[0], g2[0]
        d, i = ps.ref_nn_search(x1[None], x2[None])
        out[f"{name}_nn_dist"], out[f"{name}_nn_idx"] = d[0], i[0]
    # per-item counts (the *_dyn kernel)
    x1, x2 = out["b_xyz1"], out["b_xyz2"]
    out["b_dyn_counts"] = np.asarray([101, 60], np.int32)
    out["b_dyn_match"] = ps.ref_approx_match_dyn(x1[None], x2[None], [101], [60])[0]
    np.savez_compressed(os.path.join(OUT, "pointset_ref.npz"), **out)
    print("pointset", len(out), "arrays", sum(v.nbytes for v in out.values()), "bytes")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    checkpoints()
    column()
    canyon()
    pointset()
