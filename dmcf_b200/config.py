"""YAML config contract of the reference (o3d/utils/config.py, run_pipeline.py:103-115): one file with ``dataset:``,
``model:`` and ``pipeline:`` sections; ``model:`` keys are the constructor kwargs of the class named by
``model.name``; CLI overrides of the form ``--model.key value`` are merged with int/float/bool coercion."""
from __future__ import annotations

import yaml

from . import models


def _coerce(v):
    if not isinstance(v, str):
        return v
    low = v.lower()
    if low in ("true", "false"):
        return low == "true"
    if low in ("none", "null"):
        return None
    for cast in (int, float):
        try:
            return cast(v)
        except ValueError:
            pass
    return v


def load_config(path, overrides=None):
    """Returns {'dataset':..., 'model':..., 'pipeline':...}; ``overrides`` is {'model.key.sub': value} (o3d/utils/config.py:120-138)."""
    with open(path) as fh:
        cfg = yaml.safe_load(fh)
    for sec in ("dataset", "model", "pipeline"):
        cfg.setdefault(sec, {})
        if cfg[sec] is None:
            cfg[sec] = {}
    for key, val in (overrides or {}).items():
        parts = key.lstrip("-").split(".")
        d = cfg
        for p in parts[:-1]:
            d = d.setdefault(p, {})
        d[parts[-1]] = _coerce(val)
    return cfg


def parse_cli_overrides(argv):
    """Free-form '--a.b.c v' pairs (run_pipeline.py:46-52)."""
    out, i = {}, 0
    while i < len(argv):
        if argv[i].startswith("--") and "." in argv[i] and i + 1 < len(argv):
            out[argv[i][2:]] = argv[i + 1]
            i += 2
        else:
            i += 1
    return out


def build_model(model_cfg, **extra):
    """getattr(models, cfg.model.name)(**cfg.model) (run_pipeline.py:105-114)."""
    kw = dict(model_cfg)
    kw.update(extra)
    cls = getattr(models, kw["name"], None)
    if cls is None:
        raise NotImplementedError(f"model {kw['name']!r} is outside the hot-path scope (SURVEY 2)")
    return cls(**kw)
