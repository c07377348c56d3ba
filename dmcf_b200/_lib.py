"""ctypes binding of libdmcf_b200.so (the C ABI declared in include/dmcf_b200.h).

There is no CPU fallback: if the shared library is missing or a symbol is absent this module raises, and
every op raises when handed a non-CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdmcf_b200.so")

c_i32, c_i64, c_f32, c_vp, c_sz = C.c_int32, C.c_int64, C.c_float, C.c_void_p, C.c_size_t


class DmcfError(RuntimeError):
    pass


class Grid(C.Structure):
    """struct dmcf_grid"""
    _fields_ = [("origin", c_f32 * 3), ("inv_cell", c_f32), ("dims", c_i32 * 3), ("n_points", c_i32),
                ("cell_start", c_vp), ("sorted_index", c_vp), ("sorted_pos", c_vp), ("n_points_dev", c_vp), ("points", c_vp), ("mean_occupancy", c_f32)]


class ConvDesc(C.Structure):
    """struct dmcf_conv_desc"""
    _fields_ = [("kernel_size", c_i32 * 3), ("cin", c_i32), ("cout", c_i32), ("mapping", c_i32),
                ("interpolation", c_i32), ("align_corners", c_i32), ("normalize", c_i32), ("window", c_i32),
                ("window_fac", c_f32), ("extent", c_f32), ("offset", c_f32 * 3), ("relu_input", c_i32),
                ("feat_scale", c_f32), ("ascc", c_i32), ("skip_self", c_i32), ("nbr_lo", c_i32), ("nbr_hi", c_i32),
                ("dense_cin", c_i32), ("accumulate", c_i32), ("filter_antisym", c_i32), ("n_out_dev", c_vp),
                ("block_cin", c_i32), ("block_cout", c_i32 * 2)]


# name -> (restype, argtypes); must list every symbol of include/dmcf_b200.h
SIGNATURES = {
    "dmcf_version": (c_i32, []),
    "dmcf_last_error": (C.c_char_p, []),
    "dmcf_launch_count": (c_i64, []),
    "dmcf_grid_workspace_bytes": (c_sz, [c_i64, c_i64]),
    "dmcf_grid_build": (c_i32, [c_vp, C.POINTER(Grid), c_vp, c_sz, c_vp]),
    "dmcf_frs_count": (c_i32, [C.POINTER(Grid), c_vp, c_i64, c_vp, c_f32, c_i32, c_vp, c_vp]),
    "dmcf_frs_fill": (c_i32, [C.POINTER(Grid), c_vp, c_i64, c_vp, c_f32, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "dmcf_scan_workspace_bytes": (c_sz, [c_i64]),
    "dmcf_exclusive_scan_i32_i64": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "dmcf_exclusive_scan_i32_i32": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "dmcf_cconv_forward": (c_i32, [C.POINTER(ConvDesc), c_vp, c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp,
                                   c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "dmcf_cconv_patches": (c_i32, [C.POINTER(ConvDesc), c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp,
                                   c_i64, c_vp, c_i64, c_vp]),
    "dmcf_cconv_records_bytes": (c_sz, [c_i64]),
    "dmcf_umma_probe": (c_i32, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_vp, c_vp, c_vp]),
    "dmcf_cconv_prepare": (c_i32, [C.POINTER(ConvDesc), c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp,
                                   c_vp]),
    "dmcf_set_kernel_options": (c_i32, [c_i32]),
    "dmcf_dense_forward": (c_i32, [c_vp, c_i64, c_i32, c_i64, c_vp, c_vp, c_i32, c_i32, c_vp, c_i64, c_vp]),
    "dmcf_integrate": (c_i32, [c_vp, c_vp, c_vp, C.POINTER(c_f32), c_f32, c_i64, c_vp, c_vp, c_vp]),
    "dmcf_grid_pos_mark": (c_i32, [c_vp, c_i64, c_vp, C.POINTER(c_f32), C.POINTER(c_f32), c_vp, c_f32, C.POINTER(c_i32),
                                   C.POINTER(c_i32), c_vp, c_vp, c_vp]),
    "dmcf_grid_pos_emit": (c_i32, [c_vp, c_vp, C.POINTER(c_f32), C.POINTER(c_f32), c_vp, C.POINTER(c_i32), C.POINTER(c_i32),
                                   c_vp, c_i64, c_vp, c_vp]),
    "dmcf_correct": (c_i32, [c_vp, c_vp, c_vp, c_i64, c_i32, C.POINTER(c_f32), c_f32, c_i64, c_vp, c_vp, c_vp]),
    "dmcf_rows_append": (c_i32, [c_vp, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp]),
    "dmcf_farthest_point_sample": (c_i32, [c_vp, c_i32, c_i32, c_i32, c_vp, c_vp, c_i32, c_vp]),
    "dmcf_approx_match_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "dmcf_approx_match": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_i32, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "dmcf_match_cost_workspace_bytes": (c_sz, [c_i32, c_i32]),
    "dmcf_match_cost": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp, c_vp, c_sz, c_vp]),
    "dmcf_match_cost_grad": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "dmcf_nn_distance": (c_i32, [c_vp, c_i32, c_vp, c_i32, c_vp, c_vp, c_vp]),
}

_lib = None


def load():
    """Loads the library (once) and binds every symbol; raises DmcfError if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DmcfError(
            f"{LIB_PATH} not found: build it with `python -m dmcf_b200.build` (or __graft_entry__.build()); "
            "dmcf_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.dmcf_version() < 105:
        raise DmcfError("libdmcf_b200.so is older than this package")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().dmcf_last_error()
        raise DmcfError(f"libdmcf_b200 error {rc}: {msg.decode() if msg else ''}")
