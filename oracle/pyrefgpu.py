"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/librefgpu.so -- the REFERENCE's own GpuDetector
(built by oracle/build_ref.sh from the sources under /root/reference, see oracle/ref_harness.cu).

Needs a CUDA device.  Only tests/ and bench.py's reporting legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "librefgpu.so")
_LIB = None

POINT_DT = np.dtype([("rep0", "<u4"), ("rep1", "<u4"), ("x", "<u2"), ("y", "<u2"), ("bx", "<u2"), ("by", "<u2"),
                     ("dir", "u1"), ("b2w", "u1"), ("pad", "u1", (2,))])
EXTENTS_DT = np.dtype([("min_x", "<u2"), ("min_y", "<u2"), ("max_x", "<u2"), ("max_y", "<u2"), ("start", "<u4"),
                       ("count", "<u4"), ("gx_sum", "<i4"), ("gy_sum", "<i4"), ("pxgx_plus_pygy_sum", "<i8")])
SPOINT_DT = np.dtype([("blob", "<u4"), ("theta", "<u4"), ("x", "<u2"), ("y", "<u2"), ("bx", "<u2"), ("by", "<u2"),
                      ("dir", "u1"), ("b2w", "u1"), ("pad", "u1", (2,))])
LFP_DT = np.dtype([("Mxx", "<i8"), ("Myy", "<i8"), ("Mxy", "<i8"), ("Mx", "<i8"), ("My", "<i8"), ("W", "<i8"),
                   ("blob", "<u4"), ("pad", "<u4")])
MOMENTS_DT = np.dtype([("Mx", "<i8"), ("My", "<i8"), ("W", "<i8"), ("Mxx", "<i8"), ("Myy", "<i8"), ("Mxy", "<i8"),
                       ("N", "<i4"), ("pad", "<i4")])
FITQUAD_DT = np.dtype([("blob", "<u4"), ("valid", "<i4"), ("indices", "<u4", (4,)), ("moments", MOMENTS_DT, (4,))])
CORNERS_DT = np.dtype([("corners", "<f4", (4, 2)), ("reversed_border", "<i4"), ("blob", "<u4")])


def available() -> bool:
    return os.path.exists(SO)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        L.refgpu_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.refgpu_create.restype = C.c_void_p
        L.refgpu_destroy.argtypes = [C.c_void_p]
        L.refgpu_detect.argtypes = [C.c_void_p, C.c_void_p]
        L.refgpu_time_detect.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]
        L.refgpu_time_detect.restype = C.c_double
        for name in ("gray", "decimated", "thresholded", "labels", "sizes"):
            getattr(L, "refgpu_copy_" + name).argtypes = [C.c_void_p, C.c_void_p]
        for name in ("num_points", "num_blob_pairs", "num_selected_points", "num_fit_quads"):
            getattr(L, "refgpu_" + name).argtypes = [C.c_void_p]
            getattr(L, "refgpu_" + name).restype = C.c_int
        for name in ("sorted_points", "extents", "selected_extents", "sorted_selected", "line_fit_points", "fit_quads",
                     "quad_corners", "refined_quads"):
            f = getattr(L, "refgpu_copy_" + name)
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
            f.restype = C.c_int
        L.refgpu_copy_errors.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.refgpu_copy_errors.restype = C.c_int
        _LIB = L
    return _LIB


class ReferenceGpuDetector:
    """frc971::apriltag::GpuDetector of the reference (YUYV input, quad_decimate 2, tag36h11)."""

    def __init__(self, width, height, camera=(1.0, 0.0, 1.0, 0.0), dist=(0.0, 0.0, 0.0, 0.0, 0.0),
                 min_cluster_pixels=5, refine_edges=True):
        self.width, self.height = width, height
        self.w, self.h = width // 2, height // 2
        cam = np.asarray(camera, dtype=np.float64)
        dc = np.asarray(dist, dtype=np.float64)
        self._h = lib().refgpu_create(width, height, cam.ctypes.data, dc.ctypes.data, int(min_cluster_pixels),
                                      int(bool(refine_edges)))

    def close(self):
        if self._h:
            lib().refgpu_destroy(self._h)
            self._h = None

    def Detect(self, yuyv: np.ndarray):
        yuyv = np.ascontiguousarray(yuyv, dtype=np.uint8)
        assert yuyv.size == self.width * self.height * 2
        self._keep = yuyv
        lib().refgpu_detect(self._h, yuyv.ctypes.data)

    def time_detect(self, frames, iters, warmup=3) -> float:
        """Milliseconds per synchronous Detect call (pageable H2D included, decode stubbed)."""
        frames = [np.ascontiguousarray(f, dtype=np.uint8) for f in frames]
        ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        return float(lib().refgpu_time_detect(self._h, ptrs, len(frames), iters, warmup))

    def _img(self, name, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        getattr(lib(), "refgpu_copy_" + name)(self._h, out.ctypes.data)
        return out

    def gray(self):
        return self._img("gray", np.uint8, (self.height, self.width))

    def decimated(self):
        return self._img("decimated", np.uint8, (self.h, self.w))

    def thresholded(self):
        return self._img("thresholded", np.uint8, (self.h, self.w))

    def labels(self):
        return self._img("labels", np.uint32, (self.h * self.w,))

    def sizes(self):
        return self._img("sizes", np.uint32, (self.h * self.w,))

    def _vec(self, name, dtype, count):
        out = np.zeros(max(count, 1), dtype=dtype)
        n = getattr(lib(), "refgpu_copy_" + name)(self._h, out.ctypes.data, count)
        assert n == count, (name, n, count)
        return out[:count]

    def sorted_points(self):
        return self._vec("sorted_points", POINT_DT, lib().refgpu_num_points(self._h))

    def extents(self):
        return self._vec("extents", EXTENTS_DT, lib().refgpu_num_blob_pairs(self._h))

    def selected_extents(self):
        return self._vec("selected_extents", EXTENTS_DT, lib().refgpu_num_blob_pairs(self._h))

    def sorted_selected(self):
        return self._vec("sorted_selected", SPOINT_DT, lib().refgpu_num_selected_points(self._h))

    def line_fit_points(self):
        return self._vec("line_fit_points", LFP_DT, lib().refgpu_num_selected_points(self._h))

    def errors(self):
        n = lib().refgpu_num_selected_points(self._h)
        e = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        got = lib().refgpu_copy_errors(self._h, e[0].ctypes.data, e[1].ctypes.data, n)
        assert got == n
        return e[0][:n], e[1][:n]

    def fit_quads(self):
        return self._vec("fit_quads", FITQUAD_DT, lib().refgpu_num_fit_quads(self._h))

    def quad_corners(self):
        out = np.zeros(4096, dtype=CORNERS_DT)
        n = lib().refgpu_copy_quad_corners(self._h, out.ctypes.data, len(out))
        return out[:n]

    def refined_quads(self):
        out = np.zeros(4096, dtype=CORNERS_DT)
        n = lib().refgpu_copy_refined_quads(self._h, out.ctypes.data, len(out))
        return out[:n]
