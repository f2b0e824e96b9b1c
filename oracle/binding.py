"""ctypes binding of oracle/libklt_oracle.so (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libklt_oracle.so")
MAX_LEVELS = 8


class Params(C.Structure):
    """Mirror of lego_klt_params (include/lego_klt.h)."""
    _fields_ = [("levels", C.c_int32), ("patch_lo", C.c_int32), ("patch_hi", C.c_int32),
                ("max_iters", C.c_int32), ("inverse", C.c_int32), ("has_initial", C.c_int32),
                ("kernel", C.c_int32), ("reserved", C.c_int32), ("eps", C.c_double)]


class Stats(C.Structure):
    """Mirror of lego_klt_stats (include/lego_klt.h)."""
    _fields_ = [("n_features", C.c_uint64), ("n_success", C.c_uint64), ("n_nan", C.c_uint64),
                ("n_out_of_image", C.c_uint64), ("gn_iters", C.c_uint64 * MAX_LEVELS),
                ("n_slow_path", C.c_uint64), ("n_deferred", C.c_uint64), ("defer_reason", C.c_uint64 * 4), ("ms_h2d", C.c_float), ("ms_pyramid", C.c_float),
                ("ms_solver", C.c_float), ("ms_d2h", C.c_float)]


def make_params(levels=4, patch_lo=-3, patch_hi=3, max_iters=10, inverse=False, has_initial=True,
                kernel=0, eps=1e-2) -> Params:
    return Params(levels, patch_lo, patch_hi, max_iters, int(inverse), int(has_initial), kernel, 0, eps)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("klt_oracle.cpp", "resize_u8.cpp", "klt_oracle.h")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B" if force else "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        u8p, f32p, ip = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int)
        _lib.klt_oracle_resize_half.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, u8p]
        _lib.klt_oracle_build_pyramid.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, u8p,
                                                  C.c_size_t, ip, ip]
        _lib.klt_oracle_get_pixel_value.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                                    C.c_float, C.c_float]
        _lib.klt_oracle_get_pixel_value.restype = C.c_float
        _lib.klt_oracle_ldlt2_solve.argtypes = [C.POINTER(C.c_double)] * 3
        _lib.klt_oracle_ldlt2_solve.restype = None
        _lib.klt_oracle_track.argtypes = [C.POINTER(Params), u8p, u8p, C.c_int, C.c_int, C.c_size_t,
                                          f32p, f32p, u8p, C.c_int, C.c_int, C.POINTER(Stats)]
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _check_img(img):
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    return img.shape[0], img.shape[1], img.strides[0]


def resize_half(img: np.ndarray) -> np.ndarray:
    rows, cols, step = _check_img(img)
    out = np.empty((int(rows * 0.5), int(cols * 0.5)), np.uint8)
    rc = lib().klt_oracle_resize_half(_u8(img), cols, rows, step, _u8(out))
    if rc:
        raise RuntimeError(f"klt_oracle_resize_half rc={rc}")
    return out


def build_pyramid(img: np.ndarray, levels: int):
    """list of level images, level 0 is `img` itself."""
    rows, cols, step = _check_img(img)
    cap = rows * cols
    out = np.zeros(cap, np.uint8)
    lc = (C.c_int * levels)()
    lr = (C.c_int * levels)()
    rc = lib().klt_oracle_build_pyramid(_u8(img), cols, rows, step, levels, _u8(out), cap, lc, lr)
    if rc:
        raise RuntimeError(f"klt_oracle_build_pyramid rc={rc}")
    res, off = [img], 0
    for l in range(1, levels):
        nb = lc[l] * lr[l]
        res.append(out[off:off + nb].reshape(lr[l], lc[l]).copy())
        off += nb
    return res


def get_pixel_value(img: np.ndarray, x: float, y: float) -> float:
    rows, cols, step = _check_img(img)
    buf_len = (rows - 1) * step + cols
    return float(lib().klt_oracle_get_pixel_value(_u8(img), cols, rows, step, buf_len, x, y))


def ldlt2_solve(H, b):
    Hc = (C.c_double * 4)(*np.asarray(H, np.float64).reshape(4))
    bc = (C.c_double * 2)(*np.asarray(b, np.float64).reshape(2))
    xc = (C.c_double * 2)()
    lib().klt_oracle_ldlt2_solve(Hc, bc, xc)
    return np.array([xc[0], xc[1]])


def track(img1, img2, kp1, kp2, params: Params = None, threads: int = 1):
    """LKOpticalFlow4Layer/1Layer on the CPU oracle.  Returns (kp2_out, success(uint8), Stats)."""
    params = params or make_params()
    rows, cols, step = _check_img(img1)
    assert _check_img(img2) == (rows, cols, step)
    kp1 = np.ascontiguousarray(kp1, np.float32).reshape(-1, 2)
    out = np.ascontiguousarray(kp2, np.float32).reshape(-1, 2).copy()
    n = kp1.shape[0]
    assert out.shape[0] == n
    succ = np.zeros(max(n, 1), np.uint8)
    st = Stats()
    rc = lib().klt_oracle_track(C.byref(params), _u8(img1), _u8(img2), cols, rows, step, _f32(kp1),
                                _f32(out), _u8(succ), n, threads, C.byref(st))
    if rc:
        raise RuntimeError(f"klt_oracle_track rc={rc}")
    return out, succ[:n], st
