"""ctypes binding of oracle/_ref/libklt_ref*.so -- the reference's own translation unit (TEST INFRASTRUCTURE ONLY).

See oracle/build_ref.py for what the libraries are.  They are built in the build container (where /root/reference
exists) and travel to the GPU box as prebuilt files; nothing here reads /root/reference at run time unless a library
is missing and the reference is present.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build_ref

_libs: dict = {}


def available(half_patch: int = 3, pyramids: int = 4) -> bool:
    return os.path.exists(build_ref.lib_path(half_patch, pyramids)) or build_ref.reference_present()


def lib(half_patch: int = 3, pyramids: int = 4):
    key = (half_patch, pyramids)
    if key not in _libs:
        path = build_ref.lib_path(half_patch, pyramids)
        if build_ref.reference_present():
            build_ref.build()                      # no-op when up to date
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing and /root/reference absent: build it in the build container")
        l = C.CDLL(path)
        u8p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        l.klt_ref_track.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_size_t, f32p, f32p, u8p, C.c_int, C.c_int,
                                    C.c_int, C.c_int]
        l.klt_ref_track_pairs.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_size_t, f32p, f32p, u8p, C.c_int,
                                          C.c_int, C.c_int, C.c_int, C.c_int]
        assert l.klt_ref_half_patch() == half_patch and l.klt_ref_pyramids() == pyramids
        assert l.klt_ref_verbatim() == (1 if key == (3, 4) else 0)
        _libs[key] = l
    return _libs[key]


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def track(img1, img2, kp1, kp2, inverse=False, has_initial=True, layers=None, half_patch=3, pyramids=4):
    """legoslam::LKOpticalFlow4Layer (layers == pyramids, default) or LKOpticalFlow1Layer (layers == 1), run from
    the reference's own source text.  Returns (kp2_out float32 (n,2), success uint8 (n,))."""
    l = lib(half_patch, pyramids)
    assert img1.dtype == np.uint8 and img1.ndim == 2 and img1.strides[1] == 1 and img2.shape == img1.shape
    assert img2.strides == img1.strides
    rows, cols, step = img1.shape[0], img1.shape[1], img1.strides[0]
    kp1 = np.ascontiguousarray(kp1, np.float32).reshape(-1, 2)
    out = np.ascontiguousarray(kp2, np.float32).reshape(-1, 2).copy()
    n = kp1.shape[0]
    succ = np.zeros(max(n, 1), np.uint8)
    rc = l.klt_ref_track(_u8(img1), _u8(img2), cols, rows, step, _f32(kp1), _f32(out), _u8(succ), n, int(inverse),
                         int(has_initial), pyramids if layers is None else layers)
    if rc:
        raise RuntimeError(f"klt_ref_track rc={rc}")
    return out, succ[:n]


def track_pairs(imgs1, imgs2, kp1, kp2, threads, inverse=False, has_initial=True, half_patch=3, pyramids=4):
    """B independent pairs on `threads` host threads (one pair per thread at a time).  imgs: (B, rows, cols) uint8
    contiguous; kp: (B, n, 2).  Returns (kp2_out, success)."""
    l = lib(half_patch, pyramids)
    B, rows, cols = imgs1.shape
    assert imgs1.flags.c_contiguous and imgs2.flags.c_contiguous and imgs2.shape == imgs1.shape
    kp1 = np.ascontiguousarray(kp1, np.float32).reshape(B, -1, 2)
    out = np.ascontiguousarray(kp2, np.float32).reshape(B, -1, 2).copy()
    n = kp1.shape[1]
    succ = np.zeros((B, max(n, 1)), np.uint8)
    rc = l.klt_ref_track_pairs(_u8(imgs1), _u8(imgs2), B, cols, rows, cols, _f32(kp1), _f32(out), _u8(succ), n,
                               int(inverse), int(has_initial), pyramids, int(threads))
    if rc:
        raise RuntimeError(f"klt_ref_track_pairs rc={rc}")
    return out, succ[:, :n]
