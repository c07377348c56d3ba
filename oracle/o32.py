"""ctypes front end of the O32 C oracle (oracle/o32.c) + the reference step driven through it.

TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY (see oracle/o64.py for the rules and the PARITY UNPINNED note).
The layer and model glue is shared with O64 (``o64.cconv_layer`` / ``o64.ModelO64``) by swapping the two native ops,
so the CPU baseline does exactly what the reference does: one neighbour search per conv call
(utils/convolutions.py:354-358), the antisymmetric layer as two conv passes (:433-458), separate Dense/relu/add.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import o64

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_VARIANT = "libo32.so"
F32 = np.float32


def use_timed_build(flag=True):
    """bench.py's CPU arm: switch to libo32_timed.so (same source; FMA contraction allowed, register-blocked patch x filter
    product).  The parity tests keep the default build (-ffp-contract=off)."""
    global _LIB, _VARIANT
    _VARIANT = "libo32_timed.so" if flag else "libo32.so"
    _LIB = None


def build():
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return os.path.join(_HERE, "libo32.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, _VARIANT)
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "o32.c")):
            build()
        L = C.CDLL(path)
        L.o32_frs_count.restype = C.c_void_p
        L.o32_frs_count.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_void_p]
        L.o32_frs_fill.restype = None
        L.o32_frs_fill.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.o32_continuous_conv.restype = None
        L.o32_continuous_conv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_float,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.o32_dense.restype = None
        L.o32_dense.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.o32_num_threads.restype = C.c_int
        L.o32_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _f(a):
    return np.ascontiguousarray(a, dtype=F32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads():
    return int(lib().o32_num_threads())


def set_num_threads(n):
    lib().o32_set_num_threads(int(n))


def fixed_radius_search(points, queries, radius, ignore_query_point=False, return_distances=True):
    L = lib()
    points, queries = _f(points).reshape(-1, 3), _f(queries).reshape(-1, 3)
    nq = queries.shape[0]
    splits = np.zeros(nq + 1, np.int64)
    if points.shape[0] == 0 or nq == 0:
        return np.zeros(0, np.int32), splits, np.zeros(0, F32)
    h = L.o32_frs_count(_ptr(points), points.shape[0], _ptr(queries), nq, float(F32(radius)), int(ignore_query_point), _ptr(splits))
    total = int(splits[-1])
    index = np.empty(total, np.int32)
    dist = np.empty(total, F32) if return_distances else None
    L.o32_frs_fill(h, _ptr(points), _ptr(queries), nq, float(F32(radius)), int(ignore_query_point), _ptr(splits), _ptr(index), _ptr(dist))
    return index, splits, (dist if return_distances else np.zeros(0, F32))


_MAP = {"identity": 0, "ball_to_cube_radial": 1, "ball_to_cube_volume_preserving": 2}
_INT = {"linear": 0, "linear_border": 1, "nearest_neighbor": 2}


def continuous_conv(filters, out_positions, extents, offset, inp_positions, inp_features, inp_importance, neighbors_index,
                    neighbors_importance, neighbors_row_splits, align_corners=True, coordinate_mapping="ball_to_cube_radial",
                    normalize=False, interpolation="linear"):
    L = lib()
    filters = _f(filters)
    kz, ky, kx, cin, cout = filters.shape
    out_positions, inp_positions = _f(out_positions).reshape(-1, 3), _f(inp_positions).reshape(-1, 3)
    inp_features = _f(inp_features).reshape(inp_positions.shape[0], cin)
    n_out = out_positions.shape[0]
    out = np.empty((n_out, cout), F32)
    imp = None if inp_importance is None or len(inp_importance) == 0 else _f(inp_importance)
    nimp = None if neighbors_importance is None or len(neighbors_importance) == 0 else _f(neighbors_importance)
    off = _f(offset if offset is not None else (0, 0, 0))
    idx = np.ascontiguousarray(neighbors_index, np.int32)
    rs = np.ascontiguousarray(neighbors_row_splits, np.int64)
    extent = float(np.asarray(extents, F32).reshape(-1)[0])
    L.o32_continuous_conv(_ptr(filters), kz, ky, kx, cin, cout, _ptr(out_positions), n_out, extent, _ptr(off), _ptr(inp_positions),
                          _ptr(inp_features), _ptr(imp), _ptr(idx), _ptr(nimp), _ptr(rs), int(align_corners),
                          _MAP[coordinate_mapping], int(normalize), _INT[interpolation], _ptr(out))
    return out


def dense(x, kernel, bias):
    L = lib()
    x, kernel = _f(x), _f(kernel)
    out = np.empty((x.shape[0], kernel.shape[1]), F32)
    b = None if bias is None else _f(bias)
    L.o32_dense(_ptr(x), x.shape[0], x.shape[1], _ptr(kernel), _ptr(b), kernel.shape[1], _ptr(out))
    return out


def window(name, q, fac=1.0):
    """float32 windows (utils/tools/losses.py:8-44)."""
    w = o64.window(name, np.asarray(q, F32), fac)
    return None if w is None else w.astype(F32)


class Backend:
    """The two native ops + Dense in float32 through the C library; plugs into o64.cconv_layer / o64.ModelO64."""
    dtype = F32
    fixed_radius_search = staticmethod(fixed_radius_search)
    continuous_conv = staticmethod(continuous_conv)
    dense = staticmethod(dense)
    window = staticmethod(window)


def cconv_layer(*args, **kwargs):
    return o64.cconv_layer(*args, backend=Backend, **kwargs)


class ModelO32(o64.ModelO64):
    def __init__(self, cfg, weights):
        super().__init__(cfg, weights, backend=Backend)
