"""Host-side mirror of the reference's entry points on top of the C ABI (include/lego_klt.h).

Names, argument meaning and flag semantics follow /root/reference include/legoslam/algorithm.h:123-136:
    LKOpticalFlow1Layer(img1, img2, kp1, kp2, success, inverse=false, has_initial=true)
    LKOpticalFlow4Layer(img1, img2, kp1, kp2, success, inverse=false, has_initial=true)
Python returns (kp2, success) instead of filling reference arguments; images are uint8 numpy arrays
(cv::Mat data/cols/rows/step), keypoints are float32 (n, 2) arrays of cv::KeyPoint::pt.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import KERNEL_AUTO, KERNEL_EXACT, KERNEL_LANE, KERNEL_PATCH, KERNEL_WARP, Params, Stats  # noqa: F401


def make_params(levels=4, patch_lo=-3, patch_hi=3, max_iters=10, inverse=False, has_initial=True,
                kernel=KERNEL_AUTO, eps=1e-2) -> Params:
    """Defaults are the reference's literals (src/algorithm.cpp:40-42,113,135)."""
    return Params(levels, patch_lo, patch_hi, max_iters, int(inverse), int(has_initial), kernel, 0, eps)


class Camera(C.Structure):
    """Mirror of lego_camera (include/lego_klt.h): Camera intrinsics + pose_.matrix3x4() row-major
    (/root/reference include/legoslam/camera.h:13-24)."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("pose34", C.c_double * 12)]


def make_camera(fx, fy, cx, cy, pose34) -> Camera:
    cam = Camera(fx, fy, cx, cy)
    cam.pose34[:] = [float(v) for v in np.asarray(pose34, np.float64).reshape(12)]
    return cam


def _img_args(img: np.ndarray):
    if img.dtype != np.uint8 or img.ndim != 2 or img.strides[1] != 1:
        raise ValueError("image must be a 2-D uint8 array with unit column stride")
    return img.shape[0], img.shape[1], img.strides[0]


def pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory (lego_klt_alloc_pinned), freed with the array."""
    lib = _lib.load()
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = lib.lego_klt_alloc_pinned(max(nbytes, 1))
    if not ptr:
        raise MemoryError(f"lego_klt_alloc_pinned({nbytes}) failed")
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    weakref.finalize(buf, lib.lego_klt_free_pinned, ptr)
    return arr


class Tracker:
    """One lego_klt_ctx: a device, a stream, cached single-pair buffers."""

    def __init__(self, device: int = 0):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.lego_klt_create(device, C.byref(h)), "lego_klt_create")
        self._h = h
        self.device = device
        self._fin = weakref.finalize(self, self._lib.lego_klt_destroy, h)

    def close(self):
        self._fin()

    def set_stream(self, cuda_stream_handle: int | None):
        _lib.check(self._lib.lego_klt_set_stream(self._h, C.c_void_p(cuda_stream_handle or 0)),
                   "lego_klt_set_stream")

    def sync(self):
        _lib.check(self._lib.lego_klt_sync(self._h), "lego_klt_sync")

    def track(self, img1, img2, kp1, kp2, params: Params | None = None):
        """lego_klt_track: returns (kp2_out float32 (n,2), success uint8 (n,), Stats)."""
        params = params or make_params()
        rows, cols, step = _img_args(img1)
        if _img_args(img2) != (rows, cols, step):
            raise ValueError("img1 and img2 must have the same shape and step")
        kp1 = np.ascontiguousarray(kp1, np.float32).reshape(-1, 2)
        out = np.ascontiguousarray(kp2, np.float32).reshape(-1, 2).copy()
        n = kp1.shape[0]
        if out.shape[0] != n:
            raise ValueError("kp1 and kp2 must have the same length")
        succ = np.zeros(max(n, 1), np.uint8)
        st = Stats()
        _lib.check(self._lib.lego_klt_track(self._h, C.byref(params), img1.ctypes.data, img2.ctypes.data,
                                            cols, rows, step, kp1.ctypes.data, out.ctypes.data,
                                            succ.ctypes.data, n, C.byref(st)), "lego_klt_track")
        return out, succ[:n], st

    def LKOpticalFlow4Layer(self, img1, img2, kp1, kp2, inverse=False, has_initial=True, kernel=KERNEL_AUTO):
        out, succ, _ = self.track(img1, img2, kp1, kp2, make_params(4, inverse=inverse,
                                                                    has_initial=has_initial, kernel=kernel))
        return out, succ.astype(bool)

    def LKOpticalFlow1Layer(self, img1, img2, kp1, kp2, inverse=False, has_initial=True, kernel=KERNEL_AUTO):
        out, succ, _ = self.track(img1, img2, kp1, kp2, make_params(1, inverse=inverse,
                                                                    has_initial=has_initial, kernel=kernel))
        return out, succ.astype(bool)

    def build_pyramid(self, img: np.ndarray, levels: int = 4):
        """The pyramid part of LKOpticalFlow4Layer (src/algorithm.cpp:140-154) on the GPU."""
        rows, cols, step = _img_args(img)
        cap = rows * cols
        out = np.zeros(cap, np.uint8)
        lc = (C.c_int * levels)()
        lr = (C.c_int * levels)()
        _lib.check(self._lib.lego_klt_build_pyramid(self._h, img.ctypes.data, cols, rows, step, levels,
                                                    out.ctypes.data, cap, lc, lr), "lego_klt_build_pyramid")
        res, off = [img], 0
        for l in range(1, levels):
            nb = lc[l] * lr[l]
            res.append(out[off:off + nb].reshape(lr[l], lc[l]).copy())
            off += nb
        return res

    def triangulation(self, poses34, points_xy, sing_ratio_thr: float = 1e-3):
        """legoslam::triangulation (include/legoslam/algorithm.h:11-34) for n features seen in the same views.
        poses34: (n_views, 3, 4); points_xy: (n, n_views, 2).  Returns (pt_world (n,3) float64, ok (n,) uint8)."""
        poses = np.ascontiguousarray(poses34, np.float64).reshape(-1, 12)
        pts = np.ascontiguousarray(points_xy, np.float64).reshape(-1, poses.shape[0], 2)
        n = pts.shape[0]
        out = np.zeros((max(n, 1), 3), np.float64)
        ok = np.zeros(max(n, 1), np.uint8)
        _lib.check(self._lib.lego_klt_triangulate(self._h, poses.ctypes.data, poses.shape[0], pts.ctypes.data, n,
                                              sing_ratio_thr, out.ctypes.data, ok.ctypes.data), "lego_klt_triangulate")
        return out[:n], ok[:n]

    def triangulation_stereo(self, cam_left: Camera, cam_right: Camera, kp_left, kp_right, valid=None,
                             sing_ratio_thr: float = 1e-3):
        """pixel2camera + triangulation of tracked stereo keypoints (the body of Frontend::TriangulateNewPoints,
        src/frontend_g2o.cpp:111-155, without the caller's depth gates)."""
        kl = np.ascontiguousarray(kp_left, np.float32).reshape(-1, 2)
        kr = np.ascontiguousarray(kp_right, np.float32).reshape(-1, 2)
        n = kl.shape[0]
        v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
        out = np.zeros((max(n, 1), 3), np.float64)
        ok = np.zeros(max(n, 1), np.uint8)
        _lib.check(self._lib.lego_klt_triangulate_stereo(self._h, C.byref(cam_left), C.byref(cam_right), kl.ctypes.data,
                                                     kr.ctypes.data, None if v is None else v.ctypes.data, n,
                                                     sing_ratio_thr, out.ctypes.data, ok.ctypes.data),
                   "lego_klt_triangulate_stereo")
        return out[:n], ok[:n]

    def downscale_half(self, full: np.ndarray) -> np.ndarray:
        """cv::resize(full, out, cv::Size(), 0.5, 0.5, cv::INTER_NEAREST) of Dataset::NextFrame (src/dataset.cpp:75-77)."""
        rows, cols, step = _img_args(full)
        hr, hc = self._lib.lego_klt_half_size(rows), self._lib.lego_klt_half_size(cols)
        out = np.zeros((hr, hc), np.uint8)
        _lib.check(self._lib.lego_klt_downscale_half(self._h, full.ctypes.data, cols, rows, step, out.ctypes.data,
                                                     out.size), "lego_klt_downscale_half")
        return out

    def debug_read_level(self, level: int, rows: int):
        """Device rows of `level` after build_pyramid, aprons included: (array rows x pitch, apron_left)."""
        cap = rows * 8192
        out = np.zeros(cap, np.uint8)
        pitch, left = C.c_int(0), C.c_int(0)
        _lib.check(self._lib.lego_klt_debug_read_level(self._h, level, out.ctypes.data, cap, C.byref(pitch),
                                                       C.byref(left)), "lego_klt_debug_read_level")
        return out[:rows * pitch.value].reshape(rows, pitch.value).copy(), left.value

    def detect_features(self, img, max_corners: int, quality_level: float = 0.01, min_distance: float = 20.0, mask=None,
                        exclude=None, exclude_half: float = 10.0):
        """Frontend::DetectFeatures (src/frontend_g2o.cpp:279-297) = cv::goodFeaturesToTrack under a mask that is cleared
        around the existing features.  `img`: a uint8 array or an uploaded Image handle.  Returns (corners (n,2) float32,
        scores (n,) float32) in OpenCV's order."""
        out = np.zeros((max(max_corners, 1), 2), np.float32)
        sc = np.zeros(max(max_corners, 1), np.float32)
        n = C.c_int(0)
        ex = None if exclude is None else np.ascontiguousarray(exclude, np.float32).reshape(-1, 2)
        n_ex = 0 if ex is None else ex.shape[0]
        if isinstance(img, Image):
            if mask is not None:
                raise ValueError("image handles take an exclusion list, not a mask")
            _lib.check(self._lib.lego_klt_image_detect_features(img._h, None if not n_ex else ex.ctypes.data, n_ex, exclude_half,
                                                                max_corners, quality_level, min_distance, out.ctypes.data,
                                                                sc.ctypes.data, C.byref(n)), "lego_klt_image_detect_features")
        else:
            rows, cols, step = _img_args(img)
            if mask is not None and (mask.dtype != np.uint8 or mask.shape != img.shape or mask.strides[1] != 1):
                raise ValueError("mask must be a uint8 array of the image's shape")
            _lib.check(self._lib.lego_klt_detect_features(self._h, img.ctypes.data, cols, rows, step,
                                                          None if mask is None else mask.ctypes.data,
                                                          0 if mask is None else mask.strides[0],
                                                          None if not n_ex else ex.ctypes.data, n_ex, exclude_half, max_corners,
                                                          quality_level, min_distance, out.ctypes.data, sc.ctypes.data, C.byref(n)),
                       "lego_klt_detect_features")
        return out[:n.value].copy(), sc[:n.value].copy()

    def debug_read_eig(self, rows: int, cols: int):
        out = np.zeros((rows, cols), np.float32)
        c, r = C.c_int(0), C.c_int(0)
        _lib.check(self._lib.lego_klt_debug_read_eig(self._h, out.ctypes.data, out.size, C.byref(c), C.byref(r)),
                   "lego_klt_debug_read_eig")
        assert (r.value, c.value) == (rows, cols)
        return out

    def batch(self, batch: int, rows: int, cols: int, n_per_pair: int, levels: int = 4, step: int | None = None):
        return Batch(self, batch, rows, cols, n_per_pair, levels, step)

    def image(self, rows: int, cols: int, levels: int = 4, step: int | None = None):
        """Device-resident image with a cached pyramid (sequence mode)."""
        return Image(self, rows, cols, levels, step)

    def track_images(self, img1: "Image", img2: "Image", kp1, kp2, params: Params | None = None, want_stats: bool = True):
        """lego_klt_track_images: solver only, on two uploaded images.  want_stats=False passes a null stats pointer,
        as the C++ shim does (the reference's signature has no counters): the third result is then None."""
        params = params or make_params(img1.levels)
        kp1 = np.ascontiguousarray(kp1, np.float32).reshape(-1, 2)
        out = np.ascontiguousarray(kp2, np.float32).reshape(-1, 2).copy()
        n = kp1.shape[0]
        succ = np.zeros(max(n, 1), np.uint8)
        st = Stats() if want_stats else None
        _lib.check(self._lib.lego_klt_track_images(self._h, C.byref(params), img1._h, img2._h, kp1.ctypes.data,
                                                   out.ctypes.data, succ.ctypes.data, n,
                                                   C.byref(st) if want_stats else None),
                   "lego_klt_track_images")
        return out, succ[:n], st


def _track_frame(self, prev_left, cur_left, cur_right, kp_prev, kp_cur_guess, params=None, want_stats=False):
    """lego_klt_track_frame: temporal track prev_left -> cur_left, then (chained on the device) the stereo match
    cur_left -> cur_right of the kept features.  Returns (kp_cur, success_temporal, kp_right, success_stereo[, stats_t,
    stats_s])."""
    params = params or make_params(prev_left.levels)
    kp1 = np.ascontiguousarray(kp_prev, np.float32).reshape(-1, 2)
    cur = np.ascontiguousarray(kp_cur_guess, np.float32).reshape(-1, 2).copy()
    n = kp1.shape[0]
    right = np.zeros((max(n, 1), 2), np.float32)
    st_, ss_ = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.uint8)
    a, b = (Stats(), Stats()) if want_stats else (None, None)
    _lib.check(self._lib.lego_klt_track_frame(self._h, C.byref(params), prev_left._h, cur_left._h, cur_right._h, kp1.ctypes.data,
                                              cur.ctypes.data, st_.ctypes.data, right.ctypes.data, ss_.ctypes.data, n,
                                              C.byref(a) if want_stats else None, C.byref(b) if want_stats else None),
               "lego_klt_track_frame")
    res = (cur, st_[:n], right[:n], ss_[:n])
    return res + (a, b) if want_stats else res


Tracker.track_frame = _track_frame


class Image:
    """lego_klt_image: one uploaded image and its cached pyramid."""

    def __init__(self, tracker: Tracker, rows: int, cols: int, levels: int = 4, step: int | None = None):
        self._lib = tracker._lib
        self.tracker = tracker      # the handle points into the tracker's context: keep it alive (as Batch does)
        self.rows, self.cols, self.levels, self.step = rows, cols, levels, step or cols
        h = C.c_void_p()
        _lib.check(self._lib.lego_klt_image_create(tracker._h, cols, rows, self.step, levels, C.byref(h)),
                   "lego_klt_image_create")
        self._h = h
        self._fin = weakref.finalize(self, self._lib.lego_klt_image_destroy, h)

    def upload(self, img: np.ndarray):
        if _img_args(img) != (self.rows, self.cols, self.step):
            raise ValueError("image shape/step does not match the handle")
        _lib.check(self._lib.lego_klt_image_upload(self._h, img.ctypes.data), "lego_klt_image_upload")
        return self

    def upload_fullres(self, full: np.ndarray):
        """Dataset::NextFrame's 0.5x INTER_NEAREST halving (src/dataset.cpp:75-77) on the device, then the pyramid."""
        rows, cols, step = _img_args(full)
        _lib.check(self._lib.lego_klt_image_upload_fullres(self._h, full.ctypes.data, cols, rows, step),
                   "lego_klt_image_upload_fullres")
        return self

    def close(self):
        self._fin()


class Batch:
    """Device-resident batch of B independent image pairs (lego_klt_batch)."""

    def __init__(self, tracker: Tracker, batch: int, rows: int, cols: int, n_per_pair: int, levels: int = 4,
                 step: int | None = None):
        self._lib = tracker._lib
        self.tracker = tracker
        self.B, self.rows, self.cols, self.n, self.levels = batch, rows, cols, n_per_pair, levels
        self.step = step or cols
        h = C.c_void_p()
        _lib.check(self._lib.lego_klt_batch_create(tracker._h, batch, cols, rows, self.step, n_per_pair, levels,
                                                   C.byref(h)), "lego_klt_batch_create")
        self._h = h
        self._fin = weakref.finalize(self, self._lib.lego_klt_batch_destroy, h)

    def close(self):
        self._fin()

    def detect_features(self, image_set: int = 0, max_corners: int = 150, quality_level: float = 0.01,
                        min_distance: float = 20.0, exclude_keypoints: bool = False, exclude_half: float = 10.0):
        """lego_klt_batch_detect_features: Frontend::DetectFeatures on every image of one set of the batch (images already
        in HBM), all pairs per launch.  Returns (corners (B, max_corners, 2), counts (B,), scores (B, max_corners))."""
        out = np.zeros((self.B, max_corners, 2), np.float32)
        sc = np.zeros((self.B, max_corners), np.float32)
        cnt = np.zeros(self.B, np.int32)
        _lib.check(self._lib.lego_klt_batch_detect_features(self._h, int(image_set), int(bool(exclude_keypoints)), exclude_half,
                                                            max_corners, quality_level, min_distance, out.ctypes.data,
                                                            sc.ctypes.data, cnt.ctypes.data), "lego_klt_batch_detect_features")
        return out, cnt, sc

    def use_detected_features(self):
        """lego_klt_batch_use_detected_features: the last detection's corners become the batch's source keypoints (in HBM)."""
        _lib.check(self._lib.lego_klt_batch_use_detected_features(self._h), "lego_klt_batch_use_detected_features")

    def set_feature_counts(self, counts):
        """Ragged batch: pair b tracks its first counts[b] features only (None: all n)."""
        if counts is None:
            _lib.check(self._lib.lego_klt_batch_set_feature_counts(self._h, None), "lego_klt_batch_set_feature_counts")
            return
        c = np.ascontiguousarray(counts, np.int32)
        if c.shape != (self.B,):
            raise ValueError("counts must have one entry per pair")
        _lib.check(self._lib.lego_klt_batch_set_feature_counts(self._h, c.ctypes.data), "lego_klt_batch_set_feature_counts")

    def set_pipeline_chunks(self, chunks: int):
        _lib.check(self._lib.lego_klt_batch_set_pipeline_chunks(self._h, int(chunks)), "lego_klt_batch_set_pipeline_chunks")

    def _check_inputs(self, imgs1, imgs2, kp1, kp2):
        for a in (imgs1, imgs2):
            if a.dtype != np.uint8 or not a.flags.c_contiguous or a.size != self.B * self.rows * self.step:
                raise ValueError("images must be C-contiguous uint8 of B*rows*step bytes")
        for a in (kp1, kp2):
            if a.dtype != np.float32 or not a.flags.c_contiguous or a.size != self.B * self.n * 2:
                raise ValueError("keypoints must be C-contiguous float32 of B*n*2 values")

    def upload(self, imgs1, imgs2, kp1, kp2):
        self._check_inputs(imgs1, imgs2, kp1, kp2)
        _lib.check(self._lib.lego_klt_batch_upload(self._h, imgs1.ctypes.data, imgs2.ctypes.data,
                                                   kp1.ctypes.data, kp2.ctypes.data), "lego_klt_batch_upload")

    def run(self, params: Params | None = None):
        params = params or make_params(self.levels)
        _lib.check(self._lib.lego_klt_batch_run(self._h, C.byref(params)), "lego_klt_batch_run")

    def download(self, kp2_out=None, success=None):
        if kp2_out is None:
            kp2_out = np.empty((self.B, self.n, 2), np.float32)
        if success is None:
            success = np.empty((self.B, self.n), np.uint8)
        st = Stats()
        _lib.check(self._lib.lego_klt_batch_download(self._h, kp2_out.ctypes.data, success.ctypes.data,
                                                     C.byref(st)), "lego_klt_batch_download")
        return kp2_out, success, st

    def triangulate(self, cam_left: Camera, cam_right: Camera, sing_ratio_thr: float = 1e-3, out=None, ok=None):
        """lego_klt_batch_triangulate: world points of the tracked pairs, keypoints taken where they lie in HBM."""
        nt = self.B * self.n
        out = np.zeros((self.B, self.n, 3), np.float64) if out is None else out
        ok = np.zeros((self.B, self.n), np.uint8) if ok is None else ok
        _lib.check(self._lib.lego_klt_batch_triangulate(self._h, C.byref(cam_left), C.byref(cam_right),
                                                           sing_ratio_thr, out.ctypes.data, ok.ctypes.data),
                   "lego_klt_batch_triangulate")
        return out, ok

    def timings(self, last_n: int):
        """(ms_pyramid, ms_solver) averaged over the last `last_n` runs (CUDA events around each launch)."""
        a, b = C.c_float(), C.c_float()
        _lib.check(self._lib.lego_klt_batch_timings(self._h, last_n, C.byref(a), C.byref(b)), "lego_klt_batch_timings")
        return a.value, b.value

    def track_begin(self, imgs1, imgs2, kp1, kp2_inout, success, params: Params | None = None):
        """lego_klt_track_batched_begin: enqueue the whole call and return; the buffers belong to the call until track_end."""
        self._check_inputs(imgs1, imgs2, kp1, kp2_inout)
        params = params or make_params(self.levels)
        self._pending = (imgs1, imgs2, kp1, kp2_inout, success, params)   # keep the buffers alive
        _lib.check(self._lib.lego_klt_track_batched_begin(self._h, C.byref(params), imgs1.ctypes.data, imgs2.ctypes.data,
                                                          kp1.ctypes.data, kp2_inout.ctypes.data, success.ctypes.data),
                   "lego_klt_track_batched_begin")

    def track_end(self):
        """lego_klt_track_batched_end: wait for the call begun by track_begin; returns its Stats."""
        st = Stats()
        _lib.check(self._lib.lego_klt_track_batched_end(self._h, C.byref(st)), "lego_klt_track_batched_end")
        self._pending = None
        return st

    def track(self, imgs1, imgs2, kp1, kp2_inout, success, params: Params | None = None):
        """lego_klt_track_batched: H2D + pyramids + solver + D2H; kp2_inout is overwritten."""
        self._check_inputs(imgs1, imgs2, kp1, kp2_inout)
        params = params or make_params(self.levels)
        st = Stats()
        _lib.check(self._lib.lego_klt_track_batched(self._h, C.byref(params), imgs1.ctypes.data,
                                                    imgs2.ctypes.data, kp1.ctypes.data, kp2_inout.ctypes.data,
                                                    success.ctypes.data, C.byref(st)), "lego_klt_track_batched")
        return st


class MultiTracker:
    """lego_klt_multi: one process driving several devices, B pairs cut into contiguous blocks (one per device)."""

    def __init__(self, devices, batch: int, rows: int, cols: int, n_per_pair: int, levels: int = 4, step: int | None = None):
        self._lib = _lib.load()
        self.B, self.rows, self.cols, self.n, self.levels, self.step = batch, rows, cols, n_per_pair, levels, step or cols
        dev = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        _lib.check(self._lib.lego_klt_multi_create(dev, len(devices), batch, cols, rows, self.step, n_per_pair, levels,
                                                   C.byref(h)), "lego_klt_multi_create")
        self._h = h
        self._fin = weakref.finalize(self, self._lib.lego_klt_multi_destroy, h)

    def close(self):
        self._fin()

    def shards(self):
        """[(device, first_pair, n_pairs)] of the block partition."""
        out, i = [], 0
        while True:
            d, f, c = C.c_int(), C.c_int(), C.c_int()
            n = self._lib.lego_klt_multi_shard(self._h, i, C.byref(d), C.byref(f), C.byref(c))
            if n < 0:
                raise _lib.KltError(n, "lego_klt_multi_shard")
            out.append((d.value, f.value, c.value))
            i += 1
            if i >= n:
                return out

    def set_feature_counts(self, counts):
        c = None if counts is None else np.ascontiguousarray(counts, np.int32)
        _lib.check(self._lib.lego_klt_multi_set_feature_counts(self._h, None if c is None else c.ctypes.data),
                   "lego_klt_multi_set_feature_counts")

    def set_schedule(self, block_pairs: int):
        """lego_klt_multi_set_schedule: 0 = static contiguous blocks, > 0 = devices pull blocks of that many pairs."""
        _lib.check(self._lib.lego_klt_multi_set_schedule(self._h, int(block_pairs)), "lego_klt_multi_set_schedule")

    def last_distribution(self):
        """Pairs each device tracked in the last call (lego_klt_multi_last_distribution)."""
        out = (C.c_int * 64)()
        n = self._lib.lego_klt_multi_last_distribution(self._h, out, 64)
        if n < 0:
            raise _lib.KltError(n, "lego_klt_multi_last_distribution")
        return [int(out[i]) for i in range(n)]

    def track(self, imgs1, imgs2, kp1, kp2_inout, success, params: Params | None = None):
        """lego_klt_multi_track: every block through lego_klt_track_batched on its device, concurrently."""
        params = params or make_params(self.levels)
        st = Stats()
        _lib.check(self._lib.lego_klt_multi_track(self._h, C.byref(params), imgs1.ctypes.data, imgs2.ctypes.data,
                                                  kp1.ctypes.data, kp2_inout.ctypes.data, success.ctypes.data,
                                                  C.byref(st)), "lego_klt_multi_track")
        return st


def kernel_launches() -> int:
    """Kernels this library has launched in this process so far (lego_klt_kernel_launches)."""
    return int(_lib.load().lego_klt_kernel_launches())


_default = {}


def _default_tracker(device: int = 0) -> Tracker:
    if device not in _default:
        _default[device] = Tracker(device)
    return _default[device]


def LKOpticalFlow4Layer(img1, img2, kp1, kp2, inverse=False, has_initial=True, device=0):
    """GPU replacement of legoslam::LKOpticalFlow4Layer; returns (kp2, success[bool])."""
    return _default_tracker(device).LKOpticalFlow4Layer(img1, img2, kp1, kp2, inverse, has_initial)


def LKOpticalFlow1Layer(img1, img2, kp1, kp2, inverse=False, has_initial=True, device=0):
    """GPU replacement of legoslam::LKOpticalFlow1Layer; returns (kp2, success[bool])."""
    return _default_tracker(device).LKOpticalFlow1Layer(img1, img2, kp1, kp2, inverse, has_initial)
