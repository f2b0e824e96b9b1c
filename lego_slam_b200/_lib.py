"""ctypes loader of liblego_klt.so -- the C ABI declared in include/lego_klt.h.

The product path has no CPU fallback: if the library is missing or no B200 is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LEGO_KLT_LIB") or os.path.join(HERE, "liblego_klt.so")
MAX_LEVELS = 8

# names every build must export (tests/test_abi.py checks them against include/lego_klt.h)
EXPORTS = [
    "lego_klt_abi_version", "lego_klt_last_error", "lego_klt_default_params", "lego_klt_device_count",
    "lego_klt_create", "lego_klt_destroy", "lego_klt_set_stream", "lego_klt_track",
    "lego_klt_build_pyramid", "lego_klt_debug_read_level", "lego_klt_image_upload_fullres", "lego_klt_half_size",
    "lego_klt_downscale_half", "lego_klt_triangulate", "lego_klt_triangulate_stereo",
    "lego_klt_batch_triangulate", "lego_klt_batch_create", "lego_klt_batch_destroy", "lego_klt_batch_upload",
    "lego_klt_batch_run", "lego_klt_batch_download", "lego_klt_batch_timings", "lego_klt_track_batched",
    "lego_klt_batch_device_ptrs", "lego_klt_sync", "lego_klt_alloc_pinned", "lego_klt_free_pinned",
    "lego_klt_image_create", "lego_klt_image_destroy", "lego_klt_image_upload", "lego_klt_track_images",
    "lego_klt_kernel_launches", "lego_klt_batch_set_feature_counts", "lego_klt_batch_set_pipeline_chunks",
    "lego_klt_multi_create", "lego_klt_multi_destroy", "lego_klt_multi_shard", "lego_klt_multi_set_feature_counts",
    "lego_klt_multi_track", "lego_klt_multi_set_schedule", "lego_klt_multi_last_distribution", "lego_klt_detect_features", "lego_klt_image_detect_features", "lego_klt_batch_detect_features", "lego_klt_batch_use_detected_features", "lego_klt_debug_read_eig",
    "lego_klt_track_frame", "lego_klt_track_batched_begin", "lego_klt_track_batched_end",
]


class Params(C.Structure):
    """lego_klt_params"""
    _fields_ = [("levels", C.c_int32), ("patch_lo", C.c_int32), ("patch_hi", C.c_int32),
                ("max_iters", C.c_int32), ("inverse", C.c_int32), ("has_initial", C.c_int32),
                ("kernel", C.c_int32), ("reserved", C.c_int32), ("eps", C.c_double)]


class Stats(C.Structure):
    """lego_klt_stats"""
    _fields_ = [("n_features", C.c_uint64), ("n_success", C.c_uint64), ("n_nan", C.c_uint64),
                ("n_out_of_image", C.c_uint64), ("gn_iters", C.c_uint64 * MAX_LEVELS),
                ("n_slow_path", C.c_uint64), ("n_deferred", C.c_uint64), ("defer_reason", C.c_uint64 * 4), ("ms_h2d", C.c_float), ("ms_pyramid", C.c_float),
                ("ms_solver", C.c_float), ("ms_d2h", C.c_float)]


KERNEL_AUTO, KERNEL_EXACT, KERNEL_WARP, KERNEL_LANE, KERNEL_PATCH = 0, 1, 2, 3, 4

_lib = None


class KltError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: rc={code}: {last_error()}")
        self.code = code


def load():
    """Loads the shared library; raises if it has not been built (python -m lego_slam_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `python lego_slam_b200/build.py` "
            "(there is no CPU fallback for the KLT path)")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, f32p, ip = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int)
    pp, sp = C.POINTER(Params), C.POINTER(Stats)
    lib.lego_klt_abi_version.restype = C.c_int
    lib.lego_klt_last_error.restype = C.c_char_p
    lib.lego_klt_default_params.argtypes = [pp]
    lib.lego_klt_default_params.restype = None
    lib.lego_klt_device_count.restype = C.c_int
    lib.lego_klt_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.lego_klt_destroy.argtypes = [vp]
    lib.lego_klt_destroy.restype = None
    lib.lego_klt_set_stream.argtypes = [vp, vp]
    lib.lego_klt_sync.argtypes = [vp]
    lib.lego_klt_track.argtypes = [vp, pp, vp, vp, C.c_int, C.c_int, C.c_size_t, vp, vp, vp, C.c_int, sp]
    lib.lego_klt_build_pyramid.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, C.c_int, vp, C.c_size_t, ip, ip]
    dp = C.POINTER(C.c_double)
    lib.lego_klt_triangulate.argtypes = [vp, vp, C.c_int, vp, C.c_int, C.c_double, vp, vp]
    lib.lego_klt_triangulate_stereo.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int, C.c_double, vp, vp]
    lib.lego_klt_batch_triangulate.argtypes = [vp, vp, vp, C.c_double, vp, vp]
    lib.lego_klt_image_upload_fullres.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t]
    lib.lego_klt_half_size.argtypes = [C.c_int]
    lib.lego_klt_downscale_half.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t]
    lib.lego_klt_debug_read_level.argtypes = [vp, C.c_int, vp, C.c_size_t, ip, ip]
    lib.lego_klt_batch_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                          C.POINTER(vp)]
    lib.lego_klt_batch_destroy.argtypes = [vp]
    lib.lego_klt_batch_destroy.restype = None
    lib.lego_klt_batch_upload.argtypes = [vp, vp, vp, vp, vp]
    lib.lego_klt_batch_run.argtypes = [vp, pp]
    lib.lego_klt_batch_download.argtypes = [vp, vp, vp, sp]
    lib.lego_klt_batch_timings.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.lego_klt_track_batched.argtypes = [vp, pp, vp, vp, vp, vp, vp, sp]
    lib.lego_klt_batch_device_ptrs.argtypes = [vp] + [C.POINTER(vp)] * 6
    lib.lego_klt_image_create.argtypes = [vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(vp)]
    lib.lego_klt_image_destroy.argtypes = [vp]
    lib.lego_klt_image_destroy.restype = None
    lib.lego_klt_image_upload.argtypes = [vp, vp]
    lib.lego_klt_track_images.argtypes = [vp, pp, vp, vp, vp, vp, vp, C.c_int, sp]
    lib.lego_klt_alloc_pinned.argtypes = [C.c_size_t]
    lib.lego_klt_alloc_pinned.restype = vp
    lib.lego_klt_free_pinned.argtypes = [vp]
    lib.lego_klt_free_pinned.restype = None
    lib.lego_klt_kernel_launches.restype = C.c_longlong
    lib.lego_klt_batch_set_feature_counts.argtypes = [vp, vp]
    lib.lego_klt_batch_set_pipeline_chunks.argtypes = [vp, C.c_int]
    lib.lego_klt_multi_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int,
                                          C.POINTER(vp)]
    lib.lego_klt_multi_destroy.argtypes = [vp]
    lib.lego_klt_multi_destroy.restype = None
    lib.lego_klt_multi_shard.argtypes = [vp, C.c_int, ip, ip, ip]
    lib.lego_klt_multi_set_feature_counts.argtypes = [vp, vp]
    lib.lego_klt_multi_track.argtypes = [vp, pp, vp, vp, vp, vp, vp, sp]
    lib.lego_klt_multi_set_schedule.argtypes = [vp, C.c_int]
    lib.lego_klt_batch_detect_features.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_int, C.c_double, C.c_double, vp, vp, vp]
    lib.lego_klt_multi_last_distribution.argtypes = [vp, ip, C.c_int]
    lib.lego_klt_detect_features.argtypes = [vp, vp, C.c_int, C.c_int, C.c_size_t, vp, C.c_size_t, vp, C.c_int, C.c_float,
                                             C.c_int, C.c_double, C.c_double, vp, vp, ip]
    lib.lego_klt_image_detect_features.argtypes = [vp, vp, C.c_int, C.c_float, C.c_int, C.c_double, C.c_double, vp, vp, ip]
    lib.lego_klt_debug_read_eig.argtypes = [vp, vp, C.c_size_t, ip, ip]
    lib.lego_klt_track_batched_begin.argtypes = [vp, pp, vp, vp, vp, vp, vp]
    lib.lego_klt_track_batched_end.argtypes = [vp, sp]
    lib.lego_klt_track_frame.argtypes = [vp, pp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, sp, sp]
    _lib = lib
    return lib


def last_error() -> str:
    return load().lego_klt_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise KltError(rc, what)
