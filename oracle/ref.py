"""ctypes binding of oracle/_ref/libjsref.so = the reference's own src/cuda compiled unmodified for
sm_100a (oracle/ref_build/Makefile).  TEST INFRASTRUCTURE ONLY; needs a GPU.  Used to pin the CPU
oracle (tools/make_golden.py), by GPU differential tests, and by bench.py to time the reference's
kernels beside ours.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libjsref.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.jsref_create.restype = vp
        L.jsref_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_float] + [C.c_int] * 9
        L.jsref_destroy.argtypes = [vp]
        L.jsref_max_kp.argtypes = [vp]
        L.jsref_extract.argtypes = [vp, vp, vp, vp]
        L.jsref_level_dims.argtypes = [vp, vp, vp]
        for n in ("jsref_level_image", "jsref_level_blur", "jsref_level_score"):
            getattr(L, n).argtypes = [vp, C.c_int, vp]
        L.jsref_level_keypoints.argtypes = [vp] * 7
        L.jsref_tables.argtypes = [vp] * 6
        L.jsref_stereo_match.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.c_float, vp, vp]
        L.jsref_time_pairs.restype = C.c_double
        L.jsref_time_pairs.argtypes = [vp, vp, vp, vp, C.c_float, C.c_float, C.c_int, C.c_int]
        f, ci = C.c_float, C.c_int
        L.jsref_project_points.argtypes = [ci] + [vp] * 5 + [f] * 8 + [vp] * 4
        L.jsref_hamming_pairs.argtypes = [ci] + [vp] * 5
        L.jsref_in_frustum.argtypes = [ci] + [vp] * 12 + [f] * 4 + [ci] * 5 + [f] * 2 + [vp] * 6
        _lib = L
    return _lib


class RefEye:
    def __init__(self, height, width, n_levels=8, scale_factor=1.2, fast_n_min=9, fast_n_max=14, th_fast_min=7,
                 th_fast_max=20, tile_h=46, tile_w=46, fixed_multi_scale_tile_size=0, apply_nms_ms=0,
                 nms_ms_mode_gpu=1):
        self.H, self.W, self.L = height, width, n_levels
        self._h = lib().jsref_create(height, width, n_levels, scale_factor, fast_n_min, fast_n_max, th_fast_min,
                                     th_fast_max, tile_h, tile_w, int(fixed_multi_scale_tile_size),
                                     int(apply_nms_ms), int(nms_ms_mode_gpu))
        self.max_kp = lib().jsref_max_kp(self._h)
        hh = np.zeros(n_levels, np.int32)
        ww = np.zeros(n_levels, np.int32)
        lib().jsref_level_dims(self._h, hh.ctypes.data, ww.ctypes.data)
        self.h, self.w = hh, ww

    def close(self):
        if self._h:
            lib().jsref_destroy(self._h)
            self._h = None

    def extract(self, image):
        img = np.ascontiguousarray(image, np.uint8)
        assert img.shape == (self.H, self.W)
        kps = np.zeros(6 * self.max_kp, np.int32)
        desc = np.zeros(32 * self.max_kp, np.uint8)
        n = lib().jsref_extract(self._h, img.ctypes.data, kps.ctypes.data, desc.ctypes.data)
        return kps[:6 * n].reshape(6, n).copy(), desc[:32 * n].reshape(n, 32).copy()

    def level_image(self, l):
        o = np.zeros((self.h[l], self.w[l]), np.uint8); lib().jsref_level_image(self._h, l, o.ctypes.data); return o

    def level_blur(self, l):
        o = np.zeros((self.h[l], self.w[l]), np.uint8); lib().jsref_level_blur(self._h, l, o.ctypes.data); return o

    def level_score(self, l):
        o = np.zeros((self.h[l], self.w[l]), np.int32); lib().jsref_level_score(self._h, l, o.ctypes.data); return o

    def level_keypoints(self):
        M = self.max_kp
        x, y, s = (np.zeros(M, np.int32) for _ in range(3))
        a = np.zeros(M, np.float32)
        n = np.zeros(self.L, np.int32)
        off = np.zeros(self.L, np.int32)
        lib().jsref_level_keypoints(self._h, x.ctypes.data, y.ctypes.data, s.ctypes.data, a.ctypes.data,
                                    n.ctypes.data, off.ctypes.data)
        return x, y, s, a, n, off

    def tables(self):
        lut = np.zeros(0xFFFF, np.int32); umax = np.zeros(16, np.int32); g = np.zeros(49, np.float32)
        px = np.zeros(512, np.int8); py = np.zeros(512, np.int8)
        lib().jsref_tables(self._h, lut.ctypes.data, umax.ctypes.data, g.ctypes.data, px.ctypes.data, py.ctypes.data)
        return lut, umax, g, px, py


def stereo_match(left: RefEye, right: RefEye, n_left: int, mb: float, mbf: float, th_high=100, th_low=50):
    ur = np.zeros(max(n_left, 1), np.float32)
    dp = np.zeros(max(n_left, 1), np.float32)
    n = lib().jsref_stereo_match(left._h, right._h, th_high, th_low, mb, mbf, ur.ctypes.data, dp.ctypes.data)
    assert n == n_left
    return ur[:n], dp[:n]


def time_pairs(left: RefEye, right: RefEye, img_l, img_r, mb, mbf, iters, two_threads=True) -> float:
    il = np.ascontiguousarray(img_l, np.uint8)
    ir = np.ascontiguousarray(img_r, np.uint8)
    return lib().jsref_time_pairs(left._h, right._h, il.ctypes.data, ir.ctypes.data, mb, mbf, iters, int(two_threads))
