"""ctypes binding of libpct_b200.so (include/pct_b200.h).

There is no CPU fallback: if the CUDA library cannot be loaded (or built) the
import fails, and every call raises when it reports an error.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpct_b200.so")

PCT_OK = 0
PCT_ERR_INVALID_ARGUMENT = -1
PCT_ERR_CUDA = -2
PCT_ERR_NONFINITE = -3
PCT_ERR_K_TOO_LARGE = -4
PCT_ERR_NO_DEVICE = -5

LAYOUT_ORIGINAL = 0
LAYOUT_SLICE = 1

STATUS_EXACT_PATH = 1
STATUS_FEW_NEIGHBORS = 2
STATUS_RANK_DEFICIENT = 4
STATUS_NONFINITE = 8
STATUS_UNRESOLVED = 16
STATUS_BAD_INDEX = 32

MAX_K = 128


class IndexInfo(ctypes.Structure):
    _fields_ = [
        ("num_points", c_int64),
        ("cell_size", c_float),
        ("origin", c_float * 3),
        ("extent", c_float * 3),
        ("dims", c_int32 * 3),
        ("bits_per_axis", c_int32),
        ("num_levels", c_int32),
        ("cells_level0", c_int64),
        ("device_bytes", c_int64),
        ("est_dimension", c_float),
        ("build_launches", c_int32),
    ]


class QueryStats(ctypes.Structure):
    _fields_ = [
        ("queries", c_int64),
        ("level1_retries", c_int64),
        ("exact_path", c_int64),
        ("kernel_launches", c_int64),
        ("unstaged", c_int64),
        ("unresolved", c_int64),
        ("rank_deficient", c_int64),
    ]


# name -> (restype, argtypes); mirrors include/pct_b200.h one to one
SIGNATURES = {
    "pct_version": (c_int, []),
    "pct_last_error": (c_char_p, []),
    "pct_index_build": (c_int, [c_void_p, c_int64, c_int, c_float, c_int, c_void_p, POINTER(c_void_p)]),
    "pct_index_destroy": (c_int, [c_void_p]),
    "pct_index_get_info": (c_int, [c_void_p, POINTER(IndexInfo)]),
    "pct_index_permutation": (c_int, [c_void_p, c_void_p, c_void_p]),
    "pct_index_last_stats": (c_int, [c_void_p, c_void_p, POINTER(QueryStats)]),
    "pct_knn": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "pct_knn_points": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "pct_knn_query": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "pct_curvature_points_records": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "pct_estimate_cell_size": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, POINTER(ctypes.c_float), POINTER(ctypes.c_float)]),
    "pct_index_set_slab": (c_int, [c_void_p, c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, c_void_p]),
    "pct_release_scratch": (c_int, []),
    "pct_upload": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "pct_measure_fma_peaks": (c_int, [POINTER(c_double), POINTER(c_double), c_void_p]),
    "pct_text_shape": (c_int, [c_char_p, POINTER(c_int64), POINTER(c_int64)]),
    "pct_text_load": (c_int, [c_char_p, c_int64, c_int64, c_void_p, c_int]),
    "pct_text_load_f32": (c_int, [c_char_p, c_int64, c_int64, c_void_p, c_int]),
    "pct_pca_from_neighbors": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pct_mesh_energies": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pct_ply_shape": (c_int, [c_char_p, POINTER(c_int64), POINTER(c_int64)]),
    "pct_ply_load_f32": (c_int, [c_char_p, c_int64, c_int64, c_void_p, c_int]),
    "pct_write_points_ply": (c_int, [c_char_p, c_void_p, c_int, c_int64, c_int]),
    "pct_write_curvature_ply": (c_int, [c_char_p, c_void_p, c_void_p, c_void_p, c_int64, c_int]),
    "pct_slab_select": (c_int, [c_void_p, c_int64, c_int, c_int, ctypes.c_float, ctypes.c_float, c_void_p, POINTER(c_int64), c_void_p]),
    "pct_slab_gather": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int64, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p,
                                POINTER(c_int64), c_void_p]),
    "pct_estimate_cell_size_sample": (c_int, [c_void_p, c_int64, c_int, c_int64, POINTER(ctypes.c_float), c_int, c_void_p, POINTER(ctypes.c_float)]),
    "pct_slab_bin_blocks": (c_int64, [c_int64]),
    "pct_slab_bin_count": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, POINTER(ctypes.c_float), c_void_p, POINTER(c_int64), c_void_p]),
    "pct_slab_bin_fill": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, POINTER(ctypes.c_float), c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "pct_slab_bin_fill_peers": (c_int, [c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "pct_slab_row_ids": (c_int, [c_void_p, c_int64, c_int, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p, c_void_p]),
    "pct_index_set_peers": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "pct_slab_rows": (c_int, [c_void_p, c_int64, c_int, c_int, ctypes.c_float, ctypes.c_float, c_void_p, c_void_p]),
    "pct_ball_count": (c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p, c_int, c_void_p]),
    "pct_ball_fill": (c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p, c_int64, c_void_p, c_void_p, c_int, c_void_p]),
    "pct_fit_from_neighbors": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pct_fit_from_csr": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pct_curvature_fused_knn": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pct_curvature_fused_ball": (c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "pct_curvature_fused_knn_records": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "pct_curvature_fused_ball_records": (c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p, c_int, c_void_p]),
    "pct_plane_rotate": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pct_quadric_fit": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "pct_implicit_quadric_fit": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pct_implicit_quadric_curvature": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "pct_quadric_curvature": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
    "pct_curvature_knn_host": (c_int, [c_void_p, c_int64, c_int, c_void_p, c_void_p]),
}


def _load():
    # (re)build when the library is missing or older than its sources (digest check, needs nvcc);
    # never fall back to anything else
    from . import build as _build

    try:
        _build.build()
    except _build.NvccMissing:
        # a box without the toolkit runs the library that travelled with the tree; a failed COMPILE always raises
        if not os.path.exists(LIB_PATH):
            raise
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = the library is stale or broken
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = _load()


class PctError(RuntimeError):
    pass


def last_error() -> str:
    msg = lib.pct_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    """Turn an error code into the exception the reference would have raised."""
    if rc == PCT_OK:
        return
    msg = last_error()
    if rc == PCT_ERR_NONFINITE:
        raise ValueError(msg or "Non-finite values in input points")  # ref :273-274
    if rc == PCT_ERR_K_TOO_LARGE:
        raise IndexError(msg)  # ref :640
    if rc == PCT_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    raise PctError(f"libpct_b200 error {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())
