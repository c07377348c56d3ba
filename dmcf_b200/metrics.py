"""Validation metrics of the reference's ``run_valid`` / ``run_test`` (pipelines/simulator.py:216-263), SURVEY 8f rank 2:
``distance`` / ``chamfer_distance`` / ``compute_stats`` / ``compare_dist`` / ``merge_dicts`` (utils/evaluation_helper.py:14-90).
Given NumPy arrays they run host-side with NumPy/SciPy exactly like the reference; given CUDA tensors (what ``rollout_metrics``
passes) chamfer runs on the ``dmcf_nn_distance`` kernel and the histogram KL divergence in torch ops on the device -- no
host k-d tree, no per-frame copy of the particle sets.  ``density_loss`` (utils/tools/losses.py:380-398, on the GPU through
``compute_density`` = fixed-radius search + window) and the EMD metric (``emd_loss``, utils/tools/losses.py:401-408, on the
GPU through the approx-match kernels of dmcf_b200/pointops.py)."""
from __future__ import annotations

import numpy as np
import torch

from .losses import compute_density, get_window_func


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def distance(x, y):
    """utils/evaluation_helper.py:14-16: per-particle Euclidean distance."""
    return np.linalg.norm(_np(x) - _np(y), axis=-1)


def _on_gpu(*xs):
    return any(isinstance(x, torch.Tensor) and x.is_cuda for x in xs)


def chamfer_distance(pred, gt):
    """:25-28: for every gt point the distance to the nearest pred point.  CUDA tensors: brute-force nearest neighbour on the
    device (dmcf_nn_distance, float32 like the reference's NnDistance op), result stays a device tensor."""
    if _on_gpu(pred, gt):
        from .pointops import nearest_distance
        dev = pred.device if isinstance(pred, torch.Tensor) and pred.is_cuda else gt.device
        t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(device=dev, dtype=torch.float32)
        d2, _ = nearest_distance(t(gt), t(pred))
        return torch.sqrt(d2)
    from scipy.spatial import cKDTree
    dist, _ = cKDTree(_np(pred)).query(_np(gt))
    return dist


def compute_stats(x):
    """:31-40"""
    x = _np(x)
    return {"mean": np.mean(x), "mse": np.mean(x ** 2), "var": np.var(x), "min": np.min(x), "max": np.max(x),
            "median": np.median(x), "num_particles": x.shape[0]}


def compare_dist(x, y, bin_size=25):
    """:43-72: KL divergence between the histograms of two vector sets (5..95 percentile range, ~bin_size samples per
    bin); vectorised, same bins and counts as the reference's per-sample loop."""
    if _on_gpu(x, y):
        return _compare_dist_gpu(x, y, bin_size)
    from scipy.stats import entropy
    x, y = _np(x), _np(y)
    assert x.shape == y.shape
    cnt, dim = x.shape[0], x.shape[-1]
    bin_cnt_per_dim = int((cnt // bin_size) ** (1 / dim))
    both = np.concatenate((x, y), axis=0)
    min_v, max_v = np.percentile(both, 5, axis=0), np.percentile(both, 95, axis=0)
    bin_w = (max_v - min_v + 1e-6) / bin_cnt_per_dim
    shape = (bin_cnt_per_dim + 1,) * dim

    def hist(v):
        idx = np.clip(((v - min_v) / bin_w).astype("int32"), 0, bin_cnt_per_dim)
        h = np.zeros(shape) + 1e-5
        np.add.at(h, tuple(idx.T), 1)
        return h.reshape(-1)

    return entropy(hist(x), hist(y))


def _compare_dist_gpu(x, y, bin_size=25):
    """``compare_dist`` on the device: float64 percentiles (torch.quantile interpolates linearly like np.percentile), flat bin
    index per sample, bincount, KL divergence of the normalised histograms (what scipy.stats.entropy(pk, qk) computes)."""
    dev = x.device if isinstance(x, torch.Tensor) and x.is_cuda else y.device
    t = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))).to(device=dev, dtype=torch.float64)
    x, y = t(x), t(y)
    assert x.shape == y.shape
    cnt, dim = x.shape[0], x.shape[-1]
    b = int((cnt // bin_size) ** (1 / dim))
    both = torch.cat((x, y), dim=0)
    q = torch.tensor([0.05, 0.95], dtype=torch.float64, device=dev)
    # torch.quantile is limited to 16 M elements per call: per column
    qs = torch.stack([torch.quantile(both[:, d], q) for d in range(dim)], dim=1)
    min_v, max_v = qs[0], qs[1]
    bin_w = (max_v - min_v + 1e-6) / b
    strides = torch.tensor([(b + 1) ** (dim - 1 - d) for d in range(dim)], dtype=torch.int64, device=dev)

    def hist(v):
        idx = torch.clamp(((v - min_v) / bin_w).to(torch.int32), 0, b).to(torch.int64)  # .astype('int32') truncates toward zero
        flat = (idx * strides).sum(dim=1)
        return torch.bincount(flat, minlength=(b + 1) ** dim).to(torch.float64) + 1e-5

    pk, qk = hist(x), hist(y)
    pk, qk = pk / pk.sum(), qk / qk.sum()
    return float((pk * torch.log(pk / qk)).sum())


def merge_dicts(dicts, op, start_val=0):
    """:75-90"""
    out = {}
    for d in dicts:
        for k, v in d.items():
            out[k] = op(out.get(k, start_val), v)
    return out


def density_loss(gt, pred, gt_in=None, pred_in=None, radius=0.005, eps=0.01, win=None, use_max=False, **kwargs):
    """utils/tools/losses.py:380-398 (torch tensors on the GPU)."""
    pred_dens = compute_density(pred, pred_in, radius, win=win)
    gt_dens = compute_density(gt, gt_in, radius, win=win)
    rest_dens = gt_dens.max()
    if use_max:
        return (pred_dens.max() - rest_dens).abs() / rest_dens
    return torch.relu(pred_dens - rest_dens - eps).mean()


def rollout_metrics(pos, vel, target_pos, target_vel, box, model=None, split="valid", emd=True):
    """The per-frame metric dict of run_valid (pipelines/simulator.py:216-250) for one predicted frame."""
    dev = pos.device if isinstance(pos, torch.Tensor) else "cuda"
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
    pos, vel, target_pos, target_vel, box = t(pos), t(vel), t(target_pos), t(target_vel), t(box)
    if box.shape[0] > 0:
        pos = torch.minimum(torch.maximum(pos, box.amin(dim=0)), box.amax(dim=0))
    loss = {"mse_val": float(torch.linalg.norm(target_pos - pos, dim=-1).mean()),
            "chamfer_val": float(chamfer_distance(target_pos, pos).mean())}
    if split != "train":
        loss["dens_val"] = float(density_loss(target_pos, pos, torch.cat([pos, box], 0), torch.cat([target_pos, box], 0),
                                              win=get_window_func("poly6")))
        if model is not None:
            loss["max_dens_val"] = float(density_loss(pos, target_pos, torch.cat([pos, box], 0),
                                                      torch.cat([target_pos, box], 0), radius=model.particle_radii[0],
                                                      win=get_window_func(model.window_dens), use_max=True))
        loss["chamfer_val_2"] = float(chamfer_distance(pos, target_pos).mean())
        if emd and pos.shape[0] > 0 and target_pos.shape[0] > 0:  # pipelines/simulator.py:247-249
            from .pointops import emd_loss
            loss["emd"] = float(emd_loss(target_pos.unsqueeze(0), pos.unsqueeze(0)).mean())
        loss["vel_diff_val"] = float(compare_dist(target_vel, vel))
        loss["vel_diff_val_2"] = float(compare_dist(vel, target_vel))
    return loss
