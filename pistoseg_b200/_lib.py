"""ctypes binding of libpistoseg_b200.so (C ABI: include/pistoseg_b200.h).

There is exactly one backend.  If the shared library has not been built, or there is no sm_100 device, every
entry point raises -- nothing here falls back to torch or to the CPU oracle.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# PISTOSEG_B200_LIB: alternative build of the same library (kernel A/B experiments, tools/ab.sh)
LIB_PATH = os.environ.get("PISTOSEG_B200_LIB") or os.path.join(_HERE, "libpistoseg_b200.so")

MAX_VIEWS = 16
MAX_CLASSES = 8

FUSE_LOGIT_MEAN, FUSE_PROB_MEAN = 0, 1
MASK_NONE, MASK_FILL, MASK_NEG_INF, MASK_MULTIPLY = 0, 1, 2, 3
DECIDE_SOFTMAX, DECIDE_RAW = 0, 1
IMPL_AUTO, IMPL_GENERIC, IMPL_STREAM, IMPL_FILTER2, IMPL_FILTER4, IMPL_BAND, IMPL_STATIC, IMPL_DUO, IMPL_NARROW = 0, 1, 2, 3, 4, 5, 6, 7, 8


class PistoError(RuntimeError):
    pass


class View(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("tile_stride", C.c_int64), ("h", C.c_int32), ("w", C.c_int32),
                ("xform", C.c_int32), ("reserved", C.c_int32)]


class FuseArgs(C.Structure):
    _fields_ = [("N", C.c_int32), ("C", C.c_int32), ("T_h", C.c_int32), ("T_w", C.c_int32),
                ("fuse_mode", C.c_int32), ("mask_mode", C.c_int32), ("decide_mode", C.c_int32),
                ("bg_match", C.c_int32), ("bg_label", C.c_int32), ("low_h", C.c_int32), ("low_w", C.c_int32),
                ("impl", C.c_int32),
                ("present", C.c_void_p), ("bg", C.c_void_p), ("gt", C.c_void_p), ("label_out", C.c_void_p),
                ("fused_out", C.c_void_p), ("entropy_out", C.c_void_p), ("lowres_out", C.c_void_p),
                ("conf", C.c_void_p), ("label_raw_out", C.c_void_p)]


class TilePos(C.Structure):
    _fields_ = [("y", C.c_int32), ("x", C.c_int32), ("crop_h", C.c_int32), ("crop_w", C.c_int32)]


class ResizeDesc(C.Structure):
    _fields_ = [("tile", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("reserved", C.c_int32),
                ("iy_off", C.c_int64), ("ix_off", C.c_int64), ("bg_off", C.c_int64), ("out_off", C.c_int64)]


class MosaicQuad(C.Structure):
    _fields_ = [("flip", C.c_int32), ("warp", C.c_int32), ("crop_y", C.c_int32), ("crop_x", C.c_int32),
                ("minv", C.c_double * 6)]


class MosaicPlan(C.Structure):
    _fields_ = [("split_h", C.c_int32), ("split_w", C.c_int32), ("reserved", C.c_int32 * 2), ("quad", MosaicQuad * 4)]


class MosaicCell(C.Structure):
    _fields_ = [("tile", C.c_int32), ("cy", C.c_int16), ("cx", C.c_int16)]


# every symbol include/pistoseg_b200.h declares: (name, restype, argtypes)
_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
SYMBOLS = {
    "pisto_abi_version": (_i, []),
    "pisto_last_error": (C.c_char_p, []),
    "pisto_create": (_i, [C.POINTER(_vp), _i]),
    "pisto_destroy": (_i, [_vp]),
    "pisto_launch_count": (_i64, [_vp]),
    "pisto_filter_stats": (_i, [_vp, _vp, _i]),
    "pisto_mosaic_plan_quads": (_i, [_vp, C.c_uint64, _i64, _i64, _i, _i, _i, _d, _d, _d, _d, _d, _vp, _vp]),
    "pisto_confusion_accumulate": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _vp, _vp]),
    "pisto_fuse_argmax_confusion": (_i, [_vp, C.POINTER(View), _i, C.POINTER(FuseArgs), _vp]),
    "pisto_fuse_argmax_confusion_host": (_i, [_vp, C.POINTER(View), _i, C.POINTER(FuseArgs), _i]),
    "pisto_last_pipeline_ms": (_d, [_vp]),
    "pisto_selftest_div": (_i, [_vp, _i, _vp, _vp]),
    "pisto_upsample_bilinear": (_i, [_vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "pisto_stitch_accumulate": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "pisto_canvas_normalize": (_i, [_vp, _vp, _vp, _i, _i64, _d, _vp]),
    "pisto_canvas_axpy": (_i, [_vp, _vp, _vp, _i64, _d, _vp]),
    "pisto_resize_nearest_bg": (_i, [_vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    "pisto_argmax_f64": (_i, [_vp, _vp, _i, _i64, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp]),
    "pisto_mosaic_gather": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pisto_mosaic_pack_pool": (_i, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "pisto_mosaic_gather_packed": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pisto_mosaic_bg_integral": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "pisto_canvas_resize_accumulate": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _vp, _i, _i, _i, _vp]),
    "pisto_get_background": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "pisto_mosaic_plan_cells": (_i, [_vp, C.c_uint64, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()
_handles = {}


def load():
    """Load the shared library (no device needed).  Raises PistoError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise PistoError(
                    f"{LIB_PATH} is missing: build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
                    "pistoseg_b200 has no CPU / torch fallback.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise PistoError(f"libpistoseg_b200 error {rc}: {load().pisto_last_error().decode()}")


def handle(device_index):
    """One library handle per CUDA device, created on first use."""
    lib = load()
    with _lock:
        h = _handles.get(device_index)
        if h is None:
            out = C.c_void_p()
            check(lib.pisto_create(C.byref(out), int(device_index)))
            h = out
            _handles[device_index] = h
    return h


def launch_count(device_index=0):
    h = _handles.get(device_index)
    return int(load().pisto_launch_count(h)) if h is not None else 0


def last_pipeline_ms(device_index=0):
    h = _handles.get(device_index)
    return float(load().pisto_last_pipeline_ms(h)) if h is not None else 0.0


def filter_stats(device=0, reset=True):
    """dict(multi_tiles, exact_pixels, exact_tiles) of the filtered fusion kernels since the last reset (pisto_filter_stats)."""
    buf = (C.c_ulonglong * 4)()
    check(load().pisto_filter_stats(handle(device), buf, int(bool(reset))))
    return {"multi_tiles": int(buf[1]), "exact_pixels": int(buf[2]), "exact_tiles": int(buf[3]), "mosaic_cells_exhausted": int(buf[0])}

