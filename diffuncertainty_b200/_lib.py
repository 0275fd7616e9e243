"""ctypes binding of libvalunc.so (include/valunc.h).

There is deliberately no CPU fallback anywhere in this package: if the CUDA
library cannot be loaded, or no sm_100 device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libvalunc.so")

VU_OK = 0
VU_ERR_BAD_ARG, VU_ERR_UNSUPPORTED, VU_ERR_CUDA, VU_ERR_NO_DEVICE = -1, -2, -3, -4
VU_ABI_VERSION = 4
N_UNC, N_BINS, N_EDGES, MAX_RATERS = 3, 21, 19, 8
GT_U8, GT_I64 = 0, 1
STAT_IMAGE_SUM, STAT_THRESHOLD, STAT_AREA, STAT_DICE, STAT_CALIB, STAT_NCC, STAT_PLATT_FIT = 1, 2, 4, 8, 16, 32, 64
STAT_CLASS_COUNTS = 128
N_PLATT_BINS = 256
SLAB_RENORMALIZE, SLAB_DISCRETIZE, SLAB_LOGITS = 1, 2, 4
SLAB_F32, SLAB_BF16, SLAB_F16 = 0, 1, 2
STAT_ALL_NO_GT = STAT_IMAGE_SUM | STAT_THRESHOLD | STAT_AREA

# column layout of the per-image rows (keep in sync with valunc.h; checked in tests/test_abi.py)
F64 = dict(SUM=0, THR_SUM=3, NCC_G=6, NCC_GG=7, NCC_U=8, NCC_UU=11, NCC_GU=14, BIN_SUMS=17, COLS=80)
I64 = dict(THR_COUNT=0, AREA=3, BORDER=4, NVOX=5, BIN_TOTAL=6, BIN_TRUE=69, DICE_TP=132, DICE_PRED=140,
           DICE_GT=148, COLS=156)


class Slab(C.Structure):
    _fields_ = [("data", C.c_void_p), ("P", C.c_int64), ("B", C.c_int64), ("C", C.c_int64), ("V", C.c_int64),
                ("stride_p", C.c_int64), ("stride_b", C.c_int64), ("stride_c", C.c_int64), ("stride_v", C.c_int64),
                ("member_ptrs", C.c_void_p), ("member_ptrs_host", C.c_void_p),
                ("stride_d", C.c_int64), ("draws", C.c_int32), ("flags", C.c_uint32), ("renorm_eps", C.c_float), ("dtype", C.c_int32)]


class Gt(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("R", C.c_int32),
                ("stride_b", C.c_int64), ("stride_r", C.c_int64), ("stride_v", C.c_int64),
                ("has_ignore", C.c_int32), ("ignore_index", C.c_int64)]


class Calib(C.Structure):
    _fields_ = [("a", C.c_float), ("b", C.c_float), ("edge_u", C.c_float * N_EDGES), ("mode", C.c_int32)]


class PlattFit(C.Structure):
    _fields_ = [("edge_u", C.c_float * (N_PLATT_BINS + 1))]


class MemberOut(C.Structure):
    _fields_ = [("flags", C.c_uint32), ("eps", C.c_float), ("nll_sum", C.c_void_p), ("nll_count", C.c_void_p),
                ("nll_bad", C.c_void_p), ("ged_counts", C.c_void_p)]


class FusedArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("stat_flags", C.c_uint32), ("slab", Slab),
                ("tu", C.c_void_p), ("au", C.c_void_p), ("eu", C.c_void_p), ("labels", C.c_void_p),
                ("gt", Gt), ("threshold", C.c_float * N_UNC), ("calib", Calib * N_UNC),
                ("calib_label_lut", C.c_void_p), ("stats_f64", C.c_void_p), ("stats_i64", C.c_void_p),
                ("platt_fit", C.POINTER(PlattFit)), ("platt_i64", C.c_void_p), ("platt_f64", C.c_void_p),
                ("member_labels", C.c_void_p), ("members", MemberOut), ("class_counts", C.c_void_p)]


class MapStatsArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("stat_flags", C.c_uint32), ("B", C.c_int64), ("V", C.c_int64),
                ("maps", C.c_void_p * N_UNC), ("labels", C.c_void_p), ("gt", Gt),
                ("threshold", C.c_float * N_UNC), ("calib", Calib * N_UNC), ("calib_label_lut", C.c_void_p),
                ("ncc_gt_map", C.c_void_p), ("stats_f64", C.c_void_p), ("stats_i64", C.c_void_p),
                ("platt_fit", C.POINTER(PlattFit)), ("platt_i64", C.c_void_p), ("platt_f64", C.c_void_p),
                ("class_counts", C.c_void_p), ("n_classes", C.c_int32)]


class MemberScoresArgs(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("flags", C.c_uint32), ("slab", Slab), ("gt", Gt), ("labels", C.c_void_p),
                ("eps", C.c_float), ("nll_sum", C.c_void_p), ("nll_count", C.c_void_p), ("nll_bad", C.c_void_p),
                ("ged_counts", C.c_void_p)]


class RadixState(C.Structure):
    _fields_ = [("total", C.c_int64), ("n_rank", C.c_int32), ("n_slot", C.c_int32), ("rank", C.c_int64 * 64),
                ("residual", C.c_int64 * 64), ("prefix", C.c_uint32 * 64), ("slot", C.c_int32 * 64),
                ("slot_prefix", C.c_uint32 * 64), ("key", C.c_uint32 * 64)]


MS_NLL, MS_GED = 1, 2

EXPORTS = {
    "vu_abi_version": (C.c_int, []),
    "vu_build_info": (C.c_char_p, []),
    "vu_last_error": (C.c_char_p, []),
    "vu_device_check": (C.c_int, []),
    "vu_struct_size": (C.c_int, [C.c_int]),
    "vu_fused_pass": (C.c_int, [C.POINTER(FusedArgs), C.c_void_p]),
    "vu_fused_pass_logits": (C.c_int, [C.POINTER(FusedArgs), C.c_void_p]),
    "vu_fused_members_supported": (C.c_int, [C.POINTER(FusedArgs)]),
    "vu_map_stats": (C.c_int, [C.POINTER(MapStatsArgs), C.c_void_p]),
    "vu_patch_workspace_bytes": (C.c_int64, [C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "vu_patch_max_ws": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "vu_patch_max": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vu_border_count": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "vu_platt_invert_edges_host": (C.c_int, [C.c_double, C.c_double, C.POINTER(Calib)]),
    "vu_platt_fit_edges_host": (C.c_int, [C.POINTER(PlattFit)]),
    "vu_ged_cols": (C.c_int64, [C.c_int32, C.c_int32]),
    "vu_member_scores": (C.c_int, [C.POINTER(MemberScoresArgs), C.c_void_p]),
    "vu_radix_hist": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Gt), C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vu_radix_walk": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vu_radix_hist_state": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Gt), C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vu_quantile_select": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(Gt), C.POINTER(C.c_double), C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "vu_binned_calib": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(Gt), C.POINTER(Calib), C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "vu_quantile_select_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int64, C.POINTER(Gt), C.POINTER(C.c_double),
                                           C.c_int32, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vu_binned_calib_batch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(Gt), C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vu_synth_slab": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int64,
                                C.c_float, C.c_void_p]),
    "vu_synth_gt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int32,
                              C.c_uint64, C.c_int64, C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "vu_copy_2d_async": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "vu_set_option": (C.c_int, [C.c_char_p, C.c_int64]),
    "vu_get_counter": (C.c_int64, [C.c_char_p]),
}

_lock = threading.Lock()
_lib = None


class ValuncError(RuntimeError):
    pass


def lib_path() -> str:
    return LIB_PATH


def load():
    """Load libvalunc.so (building it in-tree with nvcc first if it is missing
    or stale).  Raises if that is impossible -- there is no other code path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        from . import build as _build
        try:
            path = _build.build()
        except Exception as exc:  # nvcc missing / compile error
            if not os.path.isfile(LIB_PATH):
                raise ValuncError(
                    f"libvalunc.so is not built and could not be built ({exc}); run "
                    "`python -m diffuncertainty_b200.build`. There is no CPU fallback.") from exc
            path = LIB_PATH
        try:
            import torch  # noqa: F401  (brings libcudart.so.12 into the process: libvalunc shares torch's CUDA runtime)
        except ImportError:
            pass
        lib = C.CDLL(path)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(lib, name)  # AttributeError here = header / library drift
            fn.restype = res
            fn.argtypes = args
        if lib.vu_abi_version() != VU_ABI_VERSION:
            raise ValuncError("libvalunc ABI version mismatch; rebuild with python -m diffuncertainty_b200.build --force")
        if lib.vu_struct_size(0) != C.sizeof(FusedArgs) or lib.vu_struct_size(1) != C.sizeof(MapStatsArgs) \
                or lib.vu_struct_size(2) != C.sizeof(Calib) or lib.vu_struct_size(3) != C.sizeof(PlattFit):
            raise ValuncError("ctypes struct layout differs from valunc.h")
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    """Map a vu_status to the exception type the reference would raise."""
    if rc == VU_OK:
        return
    msg = load().vu_last_error().decode(errors="replace")
    text = f"{what}: {msg} (vu_status {rc})"
    if rc == VU_ERR_BAD_ARG:
        raise ValueError(text)
    if rc == VU_ERR_UNSUPPORTED:
        raise NotImplementedError(text)
    raise ValuncError(text)


def require_device() -> None:
    import torch
    if not torch.cuda.is_available():
        raise ValuncError("diffuncertainty_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    check(load().vu_device_check(), "vu_device_check")


def current_stream_ptr() -> int:
    import torch
    return int(torch.cuda.current_stream().cuda_stream)


def set_option(key: str, value: int) -> None:
    check(load().vu_set_option(key.encode(), int(value)), "vu_set_option")


def get_counter(key: str) -> int:
    return int(load().vu_get_counter(key.encode()))
