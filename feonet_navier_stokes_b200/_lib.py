"""ctypes binding of libfeonet_b200.so (the C ABI in include/feonet_b200.h).

There is no CPU fallback: if the shared library is missing and cannot be built, importing the
compute entry points raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

FEO_ABI_VERSION = 2
FEO_MAT_A, FEO_MAT_B1, FEO_MAT_B2, FEO_MAT_S, FEO_MAT_M = range(5)
FEO_DENSE_M, FEO_DENSE_MT, FEO_DENSE_P = range(3)

i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)


class FeoCsr(C.Structure):
    _fields_ = [("rowptr", i32p), ("col", i32p), ("val", f32p)]


class FeoOperatorDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n", C.c_int32),
        ("A", FeoCsr), ("B1", FeoCsr), ("B2", FeoCsr), ("S", FeoCsr),
        ("n_u", C.c_int32), ("idx_i", i32p), ("idx_j", i32p),
        ("ns_precond_branch", C.c_int32), ("dt", C.c_float),
        ("dense_m", f32p), ("dense_p", f32p),
    ]


class FeoOpInfo(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_u", C.c_int32), ("has_conv", C.c_int32), ("has_seq", C.c_int32),
        ("has_dense_m", C.c_int32), ("has_dense_p", C.c_int32),
        ("nnz_a", C.c_int64), ("nnz_b1", C.c_int64), ("nnz_b2", C.c_int64), ("nnz_s", C.c_int64),
        ("nnz_union", C.c_int64), ("n_tiles_fwd", C.c_int32), ("n_tiles_bwd", C.c_int32), ("max_row_nnz", C.c_int32),
        ("device_bytes", C.c_int64),
    ]


# name -> (restype, argtypes); mirrors include/feonet_b200.h one to one
_vp, _sz, _i32, _i64, _f32 = C.c_void_p, C.c_size_t, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "feo_abi_version": (C.c_int, []),
    "feo_last_error_string": (C.c_char_p, []),
    "feo_op_create": (C.c_int, [C.POINTER(FeoOperatorDesc), C.POINTER(_vp)]),
    "feo_op_destroy": (C.c_int, [_vp]),
    "feo_op_get_info": (C.c_int, [_vp, C.POINTER(FeoOpInfo)]),
    "feo_workspace_bytes": (_sz, [_vp, _i32, _i32]),
    "feo_transpose": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp]),
    "feo_transpose_gather": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _vp]),
    "feo_op_plan": (C.c_int, [_vp]),
    "feo_residual_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _sz, _vp]),
    "feo_residual_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "feo_spmm": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _i64, _i32, _f32, _i32, _vp]),
    "feo_dense_apply": (C.c_int, [_vp, _i32, _vp, _vp, _i64, _i32, _f32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "feo_seq_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _vp, _sz, _vp]),
    "feo_seq_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "feo_assemble_u_init": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "feo_sincos_forcing_grid": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "feo_sq_diff_sum": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _f32, _vp, _vp, _sz, _vp]),
    "feo_debug_tile_replay": (C.c_int, [C.POINTER(FeoOperatorDesc), _i32, _i32, _i32, f64p, f64p, f64p, i64p]),
    "feo_debug_patch_replay": (C.c_int, [C.POINTER(FeoOperatorDesc), _i32, _i32, _i32, _i32, f64p, f64p, f64p, i64p]),
    "feo_debug_lattice_replay": (C.c_int, [C.POINTER(FeoOperatorDesc), _i32, f64p, f64p, f64p, i64p]),
    "feo_debug_dense_split_replay": (C.c_int64, [f32p, _i32, _i32, f64p, f64p, f64p]),
}

_lock = threading.Lock()
_lib = None


class FeoError(RuntimeError):
    pass


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """Load (building in-tree if needed) libfeonet_b200.so. Raises if that is impossible."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if not os.path.exists(path):
            if not build_if_missing:
                raise FeoError(f"{path} is missing and there is no CPU fallback; run __graft_entry__.build()")
            path = _build.build_library()
        lib = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if lib.feo_abi_version() != FEO_ABI_VERSION:
            raise FeoError("libfeonet_b200.so ABI version mismatch; rebuild")
        _lib = lib
        return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().feo_last_error_string()
        raise FeoError(f"libfeonet_b200 error {rc}: {msg.decode() if msg else '?'}")
