"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.
The product (jetson_slam_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


class OrcConfig(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32), ("n_levels", C.c_int32), ("scale_factor", C.c_float),
        ("fast_n_min", C.c_int32), ("fast_n_max", C.c_int32), ("th_fast_min", C.c_int32), ("th_fast_max", C.c_int32),
        ("tile_h", C.c_int32), ("tile_w", C.c_int32), ("fixed_multi_scale_tile_size", C.c_int32),
        ("apply_nms_ms", C.c_int32), ("nms_ms_mode_gpu", C.c_int32),
    ]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "jsfe_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32p, u8p, f32p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.POINTER(C.c_float)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [C.POINTER(OrcConfig), C.c_void_p]
        L.orc_destroy.argtypes = [vp]
        L.orc_max_kp.argtypes = [vp]
        L.orc_n_levels.argtypes = [vp]
        L.orc_level_geometry.argtypes = [vp, i32p]
        L.orc_scales.argtypes = [vp, f32p, f32p]
        for name, rt in (("orc_lut", u8p), ("orc_umax", i32p), ("orc_gauss", f32p),
                         ("orc_pattern_x", C.POINTER(C.c_int8)), ("orc_pattern_y", C.POINTER(C.c_int8)),
                         ("orc_cell_x", i32p), ("orc_cell_y", i32p), ("orc_cell_score", i32p),
                         ("orc_n_keypoints", i32p), ("orc_kp_x", i32p), ("orc_kp_y", i32p),
                         ("orc_kp_score", i32p), ("orc_kp_angle", f32p)):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [vp]
        for name, rt in (("orc_level_image", u8p), ("orc_level_blur", u8p), ("orc_level_score", i32p)):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [vp, C.c_int]
        L.orc_column_rank.argtypes = [C.c_int, i32p]
        L.orc_extract.argtypes = [vp, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_extract.restype = C.c_int
        for st in ("fast", "cells", "nms_ms", "compact", "orient", "blur", "describe"):
            getattr(L, "orc_stage_" + st).argtypes = [vp]
        L.orc_stage_pyramid.argtypes = [vp, C.c_void_p]
        for name in ("orc_cosf", "orc_sinf"):
            getattr(L, name).restype = C.c_float
            getattr(L, name).argtypes = [C.c_float]
        L.orc_atan2f.restype = C.c_float
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_stereo_match.restype = C.c_int
        L.orc_stereo_match.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, C.c_float,
                                       C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_stereo_pair.restype = C.c_int
        L.orc_stereo_pair.argtypes = [vp, vp, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int]
        f, vpp, ci = C.c_float, C.c_void_p, C.c_int
        L.orc_logf.restype = f
        L.orc_logf.argtypes = [f]
        L.orc_project_points.argtypes = [ci] + [vpp] * 5 + [f] * 8 + [vpp] * 4
        L.orc_hamming_pairs.argtypes = [ci] + [vpp] * 5
        L.orc_in_frustum.argtypes = [ci] + [vpp] * 12 + [f] * 4 + [ci] * 5 + [f] * 2 + [vpp] * 6
        L.orc_remap_bilinear_u8.argtypes = [vpp, ci, ci, C.c_int64, vpp, vpp, ci, ci, vpp, C.c_int64]
        L.orc_cvt_gray_u8.argtypes = [vpp, C.c_int64, ci, ci, vpp]
        L.orc_assign_features_to_grid.argtypes = [ci, vpp, vpp, f, f, f, f, vpp, vpp]
        L.orc_search_by_projection.restype = ci
        L.orc_search_by_projection.argtypes = [ci] + [vpp] * 8 + [f] * 10 + [vpp, ci, ci] + [vpp] * 7 + [ci, ci] + [vpp] * 5
        _lib = L
    return _lib


def _np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape).copy()


class Oracle:
    """One eye's extractor state (mirrors orb_cuda::ORB_GPU)."""

    def __init__(self, height, width, n_levels=8, scale_factor=1.2, fast_n_min=9, fast_n_max=14,
                 th_fast_min=7, th_fast_max=20, tile_h=46, tile_w=46, fixed_multi_scale_tile_size=0,
                 apply_nms_ms=0, nms_ms_mode_gpu=1, mask=None):
        self.cfg = OrcConfig(height, width, n_levels, scale_factor, fast_n_min, fast_n_max, th_fast_min,
                             th_fast_max, tile_h, tile_w, int(fixed_multi_scale_tile_size),
                             int(apply_nms_ms), int(nms_ms_mode_gpu))
        m = None
        if mask is not None:
            self._mask = np.ascontiguousarray(mask, np.uint8)
            assert self._mask.shape == (height, width)
            m = self._mask.ctypes.data
        self._h = lib().orc_create(C.byref(self.cfg), m)
        if not self._h:
            raise ValueError("orc_create failed (bad config)")
        self.L = n_levels
        self.max_kp = lib().orc_max_kp(self._h)
        g = np.zeros(7 * n_levels, np.int32)
        lib().orc_level_geometry(self._h, g.ctypes.data_as(C.POINTER(C.c_int32)))
        g = g.reshape(n_levels, 7)
        self.h, self.w, self.tile_h, self.tile_w, self.n_tile_h, self.n_tile_w, self.level_offset = (
            g[:, i].copy() for i in range(7))
        s = np.zeros(n_levels, np.float32)
        inv = np.zeros(n_levels, np.float32)
        lib().orc_scales(self._h, s.ctypes.data_as(C.POINTER(C.c_float)), inv.ctypes.data_as(C.POINTER(C.c_float)))
        self.scale, self.inv_scale = s, inv

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # tables
    def lut(self): return _np(lib().orc_lut(self._h), (65536,), np.uint8)
    def umax(self): return _np(lib().orc_umax(self._h), (16,), np.int32)
    def gauss(self): return _np(lib().orc_gauss(self._h), (49,), np.float32)
    def pattern(self):
        return _np(lib().orc_pattern_x(self._h), (512,), np.int8), _np(lib().orc_pattern_y(self._h), (512,), np.int8)

    def extract(self, image):
        """Returns (kps[6,N] int32 with angle as f32 bits in row 3, desc[N,32] u8)."""
        img = np.ascontiguousarray(image, np.uint8)
        assert img.shape == (self.cfg.height, self.cfg.width)
        kps = np.zeros(6 * self.max_kp, np.int32)
        desc = np.zeros(32 * self.max_kp, np.uint8)
        n = lib().orc_extract(self._h, img.ctypes.data, kps.ctypes.data, desc.ctypes.data)
        return kps[:6 * n].reshape(6, n).copy(), desc[:32 * n].reshape(n, 32).copy()

    # stage-wise
    def stage(self, name, image=None):
        if name == "pyramid":
            img = np.ascontiguousarray(image, np.uint8)
            lib().orc_stage_pyramid(self._h, img.ctypes.data)
        else:
            getattr(lib(), "orc_stage_" + name)(self._h)

    def level_image(self, l): return _np(lib().orc_level_image(self._h, l), (self.h[l], self.w[l]), np.uint8)
    def level_blur(self, l): return _np(lib().orc_level_blur(self._h, l), (self.h[l], self.w[l]), np.uint8)
    def level_score(self, l): return _np(lib().orc_level_score(self._h, l), (self.h[l], self.w[l]), np.int32)
    def cells(self):
        f = lambda fn: _np(fn(self._h), (self.max_kp,), np.int32)
        return f(lib().orc_cell_x), f(lib().orc_cell_y), f(lib().orc_cell_score)
    def n_keypoints(self): return _np(lib().orc_n_keypoints(self._h), (self.L,), np.int32)
    def level_keypoints(self):
        f = lambda fn, dt: _np(fn(self._h), (self.max_kp,), dt)
        return (f(lib().orc_kp_x, np.int32), f(lib().orc_kp_y, np.int32), f(lib().orc_kp_score, np.int32),
                f(lib().orc_kp_angle, np.float32))


def column_rank(tile_w: int) -> np.ndarray:
    r = np.zeros(tile_w, np.int32)
    lib().orc_column_rank(tile_w, r.ctypes.data_as(C.POINTER(C.c_int32)))
    return r


def stereo_match(left: Oracle, right: Oracle, kps_l, desc_l, kps_r, desc_r, mb, mbf, th_high=100, th_low=50):
    """Returns (u_right[nL] f32, depth[nL] f32, best_idx_r[nL] i32, best_dist[nL] i32)."""
    kl = np.ascontiguousarray(kps_l, np.int32)
    kr = np.ascontiguousarray(kps_r, np.int32)
    dl = np.ascontiguousarray(desc_l, np.uint8)
    dr = np.ascontiguousarray(desc_r, np.uint8)
    nl, nr = kl.shape[1], kr.shape[1]
    ur = np.zeros(max(nl, 1), np.float32)
    dp = np.zeros(max(nl, 1), np.float32)
    bi = np.zeros(max(nl, 1), np.int32)
    bd = np.zeros(max(nl, 1), np.int32)
    lib().orc_stereo_match(left._h, right._h, th_high, th_low, mb, mbf, nl, kl.ctypes.data, dl.ctypes.data,
                           nr, kr.ctypes.data, dr.ctypes.data, ur.ctypes.data, dp.ctypes.data,
                           bi.ctypes.data, bd.ctypes.data)
    return ur[:nl], dp[:nl], bi[:nl], bd[:nl]


def project_points(P, Rcw, tcw, fx, fy, cx, cy, min_x, max_x, min_y, max_y):
    P = np.ascontiguousarray(P, np.float32); Rcw = np.ascontiguousarray(Rcw, np.float32); tcw = np.ascontiguousarray(tcw, np.float32)
    n = P.shape[1]
    u, v, iz = (np.zeros(n, np.float32) for _ in range(3))
    ok = np.zeros(n, np.uint8)
    lib().orc_project_points(n, P[0].ctypes.data, P[1].ctypes.data, P[2].ctypes.data, Rcw.ctypes.data, tcw.ctypes.data, fx, fy, cx, cy,
                             min_x, max_x, min_y, max_y, u.ctypes.data, v.ctypes.data, iz.ctypes.data, ok.ctypes.data)
    return u, v, iz, ok


def hamming_pairs(idx_l, idx_r, desc_l, desc_r):
    il = np.ascontiguousarray(idx_l, np.int32); ir = np.ascontiguousarray(idx_r, np.int32)
    dl = np.ascontiguousarray(desc_l, np.uint8); dr = np.ascontiguousarray(desc_r, np.uint8)
    d = np.zeros(len(il), np.int32)
    lib().orc_hamming_pairs(len(il), il.ctypes.data, ir.ctypes.data, dl.ctypes.data, dr.ctypes.data, d.ctypes.data)
    return d


def in_frustum(P, Pn, max_distance, inv_max, inv_min, Rcw, tcw, Ow, fx, fy, cx, cy, min_x, max_x, min_y, max_y, n_levels,
               log_scale_factor, view_cos_angle):
    a = lambda x: np.ascontiguousarray(x, np.float32)
    P, Pn, md, ima, imi, Rcw, tcw, Ow = map(a, (P, Pn, max_distance, inv_max, inv_min, Rcw, tcw, Ow))
    n = P.shape[1]
    iz, u, v, vc = (np.zeros(n, np.float32) for _ in range(4))
    lvl = np.zeros(n, np.int32)
    ok = np.zeros(n, np.uint8)
    lib().orc_in_frustum(n, P[0].ctypes.data, P[1].ctypes.data, P[2].ctypes.data, Pn[0].ctypes.data, Pn[1].ctypes.data,
                         Pn[2].ctypes.data, md.ctypes.data, ima.ctypes.data, imi.ctypes.data, Rcw.ctypes.data, tcw.ctypes.data,
                         Ow.ctypes.data, fx, fy, cx, cy, min_x, max_x, min_y, max_y, n_levels, log_scale_factor, view_cos_angle,
                         iz.ctypes.data, u.ctypes.data, v.ctypes.data, lvl.ctypes.data, vc.ctypes.data, ok.ctypes.data)
    return iz, u, v, lvl, vc, ok


GRID_COLS, GRID_ROWS, HISTO_LENGTH = 64, 48, 30


def assign_features_to_grid(x, y, min_x, max_x, min_y, max_y):
    """Frame::AssignFeaturesToGrid as CSR: (cell_start[64*48+1], cell_items[n]); cell = ix*48+iy."""
    x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(y, np.float32)
    winv = np.float32(GRID_COLS) / (np.float32(max_x) - np.float32(min_x))
    hinv = np.float32(GRID_ROWS) / (np.float32(max_y) - np.float32(min_y))
    start = np.zeros(GRID_COLS * GRID_ROWS + 1, np.int32)
    items = np.full(max(len(x), 1), -1, np.int32)
    lib().orc_assign_features_to_grid(len(x), x.ctypes.data, y.ctypes.data, min_x, min_y, winv, hinv, start.ctypes.data, items.ctypes.data)
    return start, items[:start[-1]]


def search_by_projection(last, cur, Rcw, tcw, fx, fy, cx, cy, min_x, max_x, min_y, max_y, mbf, th, scale_factors, level_mode,
                         th_high=100, check_orientation=True):
    """last: dict(P[3,n], octave, angle, desc[n,32]); cur: dict(x, y, octave, angle, uright, occupied, desc[m,32]).
    -> dict(nmatches, best_idx2, best_dist, rot_bin, cur_match, hist)."""
    a = lambda v, t: np.ascontiguousarray(v, t)
    P = a(last["P"], np.float32); lo = a(last["octave"], np.int32); la = a(last["angle"], np.float32); ld = a(last["desc"], np.uint8)
    cxs = a(cur["x"], np.float32); cys = a(cur["y"], np.float32); co = a(cur["octave"], np.int32); ca = a(cur["angle"], np.float32)
    cu = a(cur["uright"], np.float32); occ = a(cur["occupied"], np.uint8); cd = a(cur["desc"], np.uint8)
    Rcw = a(Rcw, np.float32); tcw = a(tcw, np.float32); sf = a(scale_factors, np.float32)
    n, m = P.shape[1], len(cxs)
    bi = np.zeros(max(n, 1), np.int32); bd = np.zeros(max(n, 1), np.int32); rb = np.zeros(max(n, 1), np.int32)
    cm = np.zeros(max(m, 1), np.int32); hist = np.zeros(HISTO_LENGTH, np.int32)
    nm = lib().orc_search_by_projection(n, P[0].ctypes.data, P[1].ctypes.data, P[2].ctypes.data, lo.ctypes.data, la.ctypes.data,
                                        ld.ctypes.data, Rcw.ctypes.data, tcw.ctypes.data, fx, fy, cx, cy, min_x, max_x, min_y, max_y,
                                        mbf, th, sf.ctypes.data, level_mode, m, cxs.ctypes.data, cys.ctypes.data, co.ctypes.data,
                                        ca.ctypes.data, cu.ctypes.data, occ.ctypes.data, cd.ctypes.data, th_high,
                                        int(check_orientation), bi.ctypes.data, bd.ctypes.data, rb.ctypes.data, cm.ctypes.data,
                                        hist.ctypes.data)
    return dict(nmatches=nm, best_idx2=bi[:n], best_dist=bd[:n], rot_bin=rb[:n], cur_match=cm[:m], hist=hist)


def remap_bilinear(src, map_x, map_y):
    """cv::remap(src, map_x, map_y, INTER_LINEAR) for one 8-bit channel (BORDER_CONSTANT 0)."""
    src = np.ascontiguousarray(src, np.uint8); mx = np.ascontiguousarray(map_x, np.float32); my = np.ascontiguousarray(map_y, np.float32)
    dst = np.zeros(mx.shape, np.uint8)
    lib().orc_remap_bilinear_u8(src.ctypes.data, src.shape[0], src.shape[1], src.shape[1], mx.ctypes.data, my.ctypes.data,
                                mx.shape[0], mx.shape[1], dst.ctypes.data, mx.shape[1])
    return dst


def cvt_gray(img, rgb=False):
    """cv::cvtColor(img, BGR2GRAY / RGB2GRAY / BGRA2GRAY / RGBA2GRAY) for 8-bit interleaved images [h, w, 3|4]."""
    img = np.ascontiguousarray(img, np.uint8)
    dst = np.zeros(img.shape[:2], np.uint8)
    lib().orc_cvt_gray_u8(img.ctypes.data, img.shape[0] * img.shape[1], img.shape[2], 2 if rgb else 0, dst.ctypes.data)
    return dst
