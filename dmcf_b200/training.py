"""One optimisation step of the reference's ``run_train`` closure (pipelines/simulator.py:316-421) on the B200 layers:
model call(s) with ``training=True`` on the layer-by-layer path, the model's loss dict weighted by ``time_w``, optional
weight decay and per-tensor gradient clipping, Adam.  The curriculum around it (warm-up roll-in, window / iteration
schedules, data loading, logging, checkpoints) is not built."""
from __future__ import annotations

import torch


def train_step(model, optimizer, samples, targets, time_w=None, w_decay=0.0, grad_clip_norm=-1.0, scheduler=None,
               loss_fn=None):
    """``samples``: list (batch) of ``[pos, vel, acc|None, None, box, box_normals]``; ``targets``: per sample a tensor
    ``[T+1, N, 3]`` of ground-truth positions (frame 0 = the sample's frame); ``time_w``: weights of the T unrolled steps.
    Returns (total loss, dict of mean loss terms)."""
    params = [p for p in model.parameters() if p.requires_grad]
    time_w = [1.0] if time_w is None else [float(w) for w in time_w]
    optimizer.zero_grad(set_to_none=True)
    total, terms = 0.0, {}
    for sample, target in zip(samples, targets):
        inputs = list(sample)
        for t, w in enumerate(time_w):
            pos, vel = model(inputs, training=True)
            ls = model.loss([pos, vel], [inputs, target[t + 1], target[t], 0], loss_fn=loss_fn)
            for k, v in ls.items():
                total = total + w * v
                terms[k] = terms.get(k, 0.0) + float(v.detach()) * w
            inputs = [pos, vel] + inputs[2:]
    denom = sum(time_w) * len(samples)
    total = total / denom
    if w_decay > 0:
        total = total + w_decay * sum((p ** 2).sum() for p in params)
    total.backward()
    if grad_clip_norm > 0:  # tf.clip_by_norm per tensor (:414-416)
        for p in params:
            if p.grad is not None:
                n = p.grad.norm()
                if n > grad_clip_norm:
                    p.grad.mul_(grad_clip_norm / n)
    optimizer.step()
    if scheduler is not None:
        scheduler.step()
    return float(total.detach()), {k: v / denom for k, v in terms.items()}
