"""Host-side binding of libjsfe.so (ctypes over the C ABI in include/jsfe.h).

`Frontend` is the batched B200 entry point (one handle = many image slots on one GPU);
`ORBExtractor` / `compute_stereo_matches` mirror the reference's per-frame interface
(Jetson_SLAM::ORBExtractor, include/ORBextractor.h:21-42; ORB_GPU::ORB_compute_stereo_match,
include/cuda/orb_gpu.hpp:218-229) with the same argument meaning and output layout, so parity
tests read like calls into the reference.

There is NO fallback: if libjsfe.so is missing or no CUDA device is present this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjsfe.so")


class JsfeError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32), ("n_levels", C.c_int32), ("scale_factor", C.c_float),
        ("fast_n_min", C.c_int32), ("fast_n_max", C.c_int32), ("th_fast_min", C.c_int32), ("th_fast_max", C.c_int32),
        ("tile_h", C.c_int32), ("tile_w", C.c_int32), ("fixed_multi_scale_tile_size", C.c_int32),
        ("apply_nms_ms", C.c_int32), ("nms_ms_mode_gpu", C.c_int32),
        ("mask", C.c_void_p), ("mask_pitch", C.c_int64), ("device_id", C.c_int32), ("max_images", C.c_int32),
    ]


class LevelInfo(C.Structure):
    _fields_ = [("height", C.c_int32), ("width", C.c_int32), ("pitch", C.c_int32), ("tile_h", C.c_int32),
                ("tile_w", C.c_int32), ("n_tile_h", C.c_int32), ("n_tile_w", C.c_int32), ("cell_offset", C.c_int32),
                ("scale", C.c_float), ("inv_scale", C.c_float)]


class SlotView(C.Structure):
    _fields_ = [("n_keypoints", C.c_void_p), ("n_per_level", C.c_void_p), ("kps", C.c_void_p), ("desc", C.c_void_p),
                ("u_right", C.c_void_p), ("depth", C.c_void_p), ("best_idx_r", C.c_void_p), ("best_dist", C.c_void_p),
                ("capacity", C.c_int32)]


class HostResults(C.Structure):
    _fields_ = [("n_keypoints", C.POINTER(C.c_int32)), ("kps", C.POINTER(C.c_int32)), ("desc", C.POINTER(C.c_uint8)),
                ("u_right", C.POINTER(C.c_float)), ("depth", C.POINTER(C.c_float)), ("capacity", C.c_int32),
                ("bytes", C.c_int64)]


class Gathered(C.Structure):
    _fields_ = [("data", C.c_void_p), ("region_stride", C.c_int64), ("world", C.c_int32), ("root", C.c_int32),
                ("n_pairs", C.c_int32), ("transport", C.c_int32)]


_lib = None


def lib():
    """Load libjsfe.so (raises if it has not been built: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JsfeError(f"{LIB_PATH} not built; run __graft_entry__.build(). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.jsfe_last_error.restype = C.c_char_p
        L.jsfe_create.argtypes = [C.POINTER(_Config), C.POINTER(vp)]
        L.jsfe_destroy.argtypes = [vp]
        L.jsfe_max_keypoints.argtypes = [vp]
        L.jsfe_num_levels.argtypes = [vp]
        L.jsfe_get_level_info.argtypes = [vp, C.c_int, C.POINTER(LevelInfo)]
        L.jsfe_set_images.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int64, C.c_int64, C.c_int, vp]
        L.jsfe_slot_image.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_int64)]
        L.jsfe_extract.argtypes = [vp, C.c_int, C.c_int, vp]
        L.jsfe_stereo_match.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, vp]
        L.jsfe_slot_view_get.argtypes = [vp, C.c_int, C.POINTER(SlotView)]
        L.jsfe_level_image.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                       C.POINTER(C.c_int64)]
        L.jsfe_pack_keypoints.argtypes = [vp, C.c_int, vp, vp, C.POINTER(C.c_int32), vp]
        L.jsfe_get_keypoints.argtypes = [vp, C.c_int, vp, vp, C.POINTER(C.c_int32), vp]
        L.jsfe_pack_keypoints_once.argtypes = [vp, C.c_int, vp, vp, C.POINTER(C.c_int32), vp]
        L.jsfe_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t, C.c_int]
        L.jsfe_host_free.argtypes = [vp]
        L.jsfe_get_stereo.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.POINTER(C.c_int32), vp]
        L.jsfe_download_results.argtypes = [vp, C.c_int, C.c_int, C.POINTER(HostResults), vp]
        L.jsfe_process_host_pairs.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.POINTER(HostResults)]
        f = C.c_float
        L.jsfe_project_points.argtypes = [C.c_int] + [vp] * 5 + [f] * 8 + [vp] * 4 + [vp]
        L.jsfe_hamming_pairs.argtypes = [C.c_int] + [vp] * 5 + [vp]
        L.jsfe_in_frustum.argtypes = [C.c_int] + [vp] * 12 + [f] * 4 + [C.c_int] * 5 + [f] * 2 + [vp] * 6 + [vp]
        L.jsfe_process_host_pairs_begin.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, f, f]
        L.jsfe_process_host_pairs_end.argtypes = [vp, C.POINTER(HostResults)]
        L.jsfe_build_frame_grid.argtypes = [C.c_int, vp, vp, f, f, f, f, vp, vp, vp]
        L.jsfe_search_by_projection.argtypes = [vp, vp]
        L.jsfe_frame_view.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp]
        i64 = C.c_int64
        L.jsfe_remap_bilinear.argtypes = [vp, C.c_int, C.c_int, i64, i64, C.c_int, vp, vp, C.c_int, C.c_int, vp, i64, i64, vp]
        L.jsfe_cvt_gray.argtypes = [vp, C.c_int, C.c_int, i64, C.c_int, C.c_int, vp, i64, vp]
        L.jsfe_debug_level_image.argtypes = [vp, C.c_int, C.c_int, vp]
        L.jsfe_debug_level_blur.argtypes = [vp, C.c_int, C.c_int, vp]
        L.jsfe_debug_cells.argtypes = [vp, C.c_int, vp, vp, vp]
        L.jsfe_debug_level_keypoints.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp]
        L.jsfe_profile_enable.argtypes = [vp, C.c_int]
        L.jsfe_profile_read.argtypes = [vp, vp, vp, C.c_int]
        L.jsfe_launch_count.restype = C.c_int64
        L.jsfe_launch_count.argtypes = [vp]
        L.jsfe_mappool_create.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
        L.jsfe_mappool_destroy.argtypes = [vp]
        L.jsfe_mappool_update.argtypes = [vp, C.c_int] + [vp] * 10 + [vp]
        L.jsfe_mappool_in_frustum.argtypes = [vp, C.c_int, vp, vp, vp, vp] + [f] * 4 + [C.c_int] * 5 + [f] * 2 + [vp] * 6 + [vp]
        L.jsfe_gather_region_bytes.restype = C.c_int64
        L.jsfe_gather_region_bytes.argtypes = [vp, C.c_int]
        L.jsfe_gather_create.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.jsfe_gather_ipc_export.argtypes = [vp, C.c_int, vp]
        L.jsfe_gather_ipc_import.argtypes = [vp, C.c_int, vp]
        L.jsfe_gather_set_peers_mapped.argtypes = [vp, C.c_int]
        L.jsfe_gather_profile.argtypes = [vp, C.c_int]
        L.jsfe_gather_stage_times.argtypes = [vp, C.POINTER(C.c_float)]
        L.jsfe_gather_begin.argtypes = [vp, C.c_int, C.c_int, vp]
        L.jsfe_gather_end.argtypes = [vp, C.POINTER(Gathered)]
        L.jsfe_gather_destroy.argtypes = [vp]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise JsfeError(f"jsfe error {rc}: {lib().jsfe_last_error().decode()}")


def _stream_ptr(stream):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):  # torch.cuda.Stream
        return C.c_void_p(stream.cuda_stream)
    return C.c_void_p(int(stream))


class HostBuffer:
    """Pinned host memory from jsfe_host_alloc as a numpy array (`.array`); write_combined=True for input images the host only
    writes (the upload is then not snooped through the CPU caches)."""

    def __init__(self, shape, dtype=np.uint8, write_combined=False):
        self._ptr = C.c_void_p()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        _check(lib().jsfe_host_alloc(C.byref(self._ptr), n, int(bool(write_combined))))
        self.array = np.frombuffer((C.c_uint8 * n).from_address(self._ptr.value), dtype=dtype).reshape(shape)
        self.write_combined = bool(write_combined)

    def close(self):
        if self._ptr:
            self.array = None
            lib().jsfe_host_free(self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Frontend:
    """One GPU, `max_images` image slots.  Pair p = slots (2p, 2p+1)."""

    def __init__(self, height, width, n_levels=8, scale_factor=1.2, fast_n_min=9, fast_n_max=14, th_fast_min=7,
                 th_fast_max=20, tile_h=30, tile_w=30, fixed_multi_scale_tile_size=0, apply_nms_ms=0,
                 nms_ms_mode_gpu=1, mask=None, device=0, max_images=2):
        self._h = None
        self._mask = None
        cfg = _Config(height, width, n_levels, scale_factor, fast_n_min, fast_n_max, th_fast_min, th_fast_max,
                      tile_h, tile_w, int(fixed_multi_scale_tile_size), int(apply_nms_ms), int(nms_ms_mode_gpu),
                      None, 0, device, max_images)
        if mask is not None:
            self._mask = np.ascontiguousarray(mask, np.uint8)
            if self._mask.shape != (height, width):
                raise ValueError("mask must be height x width")
            cfg.mask = self._mask.ctypes.data
            cfg.mask_pitch = width
        h = C.c_void_p()
        _check(lib().jsfe_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.height, self.width, self.n_levels, self.max_images = height, width, n_levels, max_images
        self.device = device
        self.max_kp = lib().jsfe_max_keypoints(self._h)
        self.levels = []
        for l in range(n_levels):
            li = LevelInfo()
            _check(lib().jsfe_get_level_info(self._h, l, C.byref(li)))
            self.levels.append(li)
        self.scale = np.array([li.scale for li in self.levels], np.float32)
        self.inv_scale = np.array([li.inv_scale for li in self.levels], np.float32)

    def close(self):
        if self._h is not None:
            lib().jsfe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- input
    def set_images(self, images, first_slot=0, stream=None):
        """images: u8 array [n, H, W] (or [H, W]) in host memory (pinned memory makes the copy asynchronous)."""
        a = np.asarray(images)
        if a.ndim == 2:
            a = a[None]
        if a.dtype != np.uint8 or a.shape[1:] != (self.height, self.width) or not a.flags.c_contiguous:
            raise ValueError("images must be C-contiguous uint8 [n, H, W]")
        _check(lib().jsfe_set_images(self._h, first_slot, a.shape[0], a.ctypes.data, self.width,
                                     self.width * self.height, 0, _stream_ptr(stream)))

    def set_images_ptr(self, ptr, n, row_pitch, image_stride, on_device, first_slot=0, stream=None):
        _check(lib().jsfe_set_images(self._h, first_slot, n, C.c_void_p(int(ptr)), row_pitch, image_stride,
                                     int(bool(on_device)), _stream_ptr(stream)))

    def slot_image(self, slot):
        p, pitch = C.c_void_p(), C.c_int64()
        _check(lib().jsfe_slot_image(self._h, slot, C.byref(p), C.byref(pitch)))
        return p.value, pitch.value

    # ---- compute (asynchronous on `stream`)
    def extract(self, first_slot=0, n=None, stream=None):
        n = self.max_images - first_slot if n is None else n
        _check(lib().jsfe_extract(self._h, first_slot, n, _stream_ptr(stream)))

    def stereo_match(self, mb, mbf, first_pair=0, n=None, th_high=100, th_low=50, stream=None):
        n = self.max_images // 2 - first_pair if n is None else n
        _check(lib().jsfe_stereo_match(self._h, first_pair, n, th_high, th_low, mb, mbf, _stream_ptr(stream)))

    # ---- results (synchronise)
    def get_keypoints(self, slot, stream=None):
        """-> (kps int32 [6, N] = x|y|score|angle_deg(f32 bits)|octave|size, desc u8 [N, 32]) in the reference layout."""
        kps = np.zeros(6 * self.max_kp, np.int32)
        desc = np.zeros(32 * self.max_kp, np.uint8)
        n = C.c_int32()
        _check(lib().jsfe_get_keypoints(self._h, slot, kps.ctypes.data, desc.ctypes.data, C.byref(n), _stream_ptr(stream)))
        n = n.value
        return kps[:6 * n].reshape(6, n).copy(), desc[:32 * n].reshape(n, 32).copy()

    def get_stereo(self, pair, stream=None):
        """-> (u_right f32[nL], depth f32[nL], best_idx_r i32[nL], best_dist i32[nL])."""
        bufs = [np.zeros(self.max_kp, dt) for dt in (np.float32, np.float32, np.int32, np.int32)]
        n = C.c_int32()
        _check(lib().jsfe_get_stereo(self._h, pair, *[b.ctypes.data for b in bufs], C.byref(n), _stream_ptr(stream)))
        return tuple(b[:n.value].copy() for b in bufs)

    def download(self, first_slot=0, n=None, stream=None):
        """Bulk D2H of n slots' result slabs into pinned staging; returns numpy views (valid until the next call)."""
        n = self.max_images - first_slot if n is None else n
        r = HostResults()
        _check(lib().jsfe_download_results(self._h, first_slot, n, C.byref(r), _stream_ptr(stream)))
        cap = r.capacity
        as_np = np.ctypeslib.as_array
        return {
            "n": as_np(r.n_keypoints, (n,)), "kps": as_np(r.kps, (n, 6, cap)), "desc": as_np(r.desc, (n, cap, 32)),
            "u_right": as_np(r.u_right, (n, cap)), "depth": as_np(r.depth, (n, cap)), "bytes": r.bytes,
        }

    def _host_results(self, r, n):
        cap = r.capacity
        as_np = np.ctypeslib.as_array
        return {"n": as_np(r.n_keypoints, (n,)), "kps": as_np(r.kps, (n, 6, cap)), "desc": as_np(r.desc, (n, cap, 32)),
                "u_right": as_np(r.u_right, (n, cap)), "depth": as_np(r.depth, (n, cap)), "bytes": r.bytes}

    @staticmethod
    def _check_images(a, height, width):
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[1:] != (height, width) or not a.flags.c_contiguous or a.shape[0] % 2:
            raise ValueError("images must be C-contiguous uint8 [2*n_pairs, H, W]")

    def process_host_pairs(self, images, mb, mbf, chunk_pairs=0, th_high=100, th_low=50):
        """End-to-end: images u8 [2*n_pairs, H, W] in host memory (L0,R0,L1,R1,...) -> dict of pinned result views.
        Upload, extract, match and download are pipelined in chunks over three CUDA streams; synchronous."""
        a = np.asarray(images)
        self._check_images(a, self.height, self.width)
        n = a.shape[0]
        r = HostResults()
        _check(lib().jsfe_process_host_pairs(self._h, n // 2, a.ctypes.data, chunk_pairs, th_high, th_low, mb, mbf, C.byref(r)))
        return self._host_results(r, n)

    def process_host_pairs_begin(self, images, mb, mbf, chunk_pairs=0, th_high=100, th_low=50):
        """Enqueue a batch and return (jsfe_process_host_pairs_begin); `images` must stay alive until process_host_pairs_end."""
        a = np.asarray(images)
        self._check_images(a, self.height, self.width)
        self._inflight = a
        _check(lib().jsfe_process_host_pairs_begin(self._h, a.shape[0] // 2, a.ctypes.data, chunk_pairs, th_high, th_low, mb, mbf))

    def process_host_pairs_end(self):
        """Wait for the batch enqueued by process_host_pairs_begin -> dict of pinned result views."""
        r = HostResults()
        _check(lib().jsfe_process_host_pairs_end(self._h, C.byref(r)))
        n = self._inflight.shape[0]
        self._inflight = None
        return self._host_results(r, n)

    def slot_view(self, slot):
        v = SlotView()
        _check(lib().jsfe_slot_view_get(self._h, slot, C.byref(v)))
        return v

    STAGES = ("k_pyramid", "k_fast_cells", "k_compact", "k_orient_desc", "k_stereo_match", "k_stereo_outlier", "k_nms_ms", "k_blur")

    def profile(self, on=True):
        _check(lib().jsfe_profile_enable(self._h, int(on)))

    def profile_read(self):
        """-> {kernel: (total_ms, launches)} since the last read (CUDA events on the launching stream)."""
        ms = np.zeros(len(self.STAGES), np.float32)
        cnt = np.zeros(len(self.STAGES), np.int64)
        _check(lib().jsfe_profile_read(self._h, ms.ctypes.data, cnt.ctypes.data, len(self.STAGES)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.STAGES)}

    def launch_count(self):
        return int(lib().jsfe_launch_count(self._h))

    # ---- stage inspection (tests)
    def level_image(self, slot, level):
        li = self.levels[level]
        out = np.zeros((li.height, li.width), np.uint8)
        _check(lib().jsfe_debug_level_image(self._h, slot, level, out.ctypes.data))
        return out

    def level_blur(self, slot, level):
        li = self.levels[level]
        out = np.zeros((li.height, li.width), np.uint8)
        _check(lib().jsfe_debug_level_blur(self._h, slot, level, out.ctypes.data))
        return out

    def cells(self, slot):
        x, y, s = (np.zeros(self.max_kp, np.int32) for _ in range(3))
        _check(lib().jsfe_debug_cells(self._h, slot, x.ctypes.data, y.ctypes.data, s.ctypes.data))
        return x, y, s

    def level_keypoints(self, slot):
        x, y, s, l = (np.zeros(self.max_kp, np.int32) for _ in range(4))
        a = np.zeros(self.max_kp, np.float32)
        _check(lib().jsfe_debug_level_keypoints(self._h, slot, x.ctypes.data, y.ctypes.data, s.ctypes.data,
                                                l.ctypes.data, a.ctypes.data))
        return x, y, s, l, a


class ORBExtractor:
    """Mirror of Jetson_SLAM::ORBExtractor (include/ORBextractor.h:21-98): same constructor argument order,
    `extract(image)` returns the keypoint SoA and descriptors the reference leaves in its SyncedMem outputs."""

    def __init__(self, im_height, im_width, scale_factor, n_levels, FAST_N_MIN, FAST_N_MAX, th_FAST_MIN, th_FAST_MAX,
                 str_mask, tile_h, tile_w, fixed_multi_scale_tile_size, apply_nms_ms, nms_ms_mode_gpu, use_gpu=True,
                 mask=None, device=0, _frontend=None, _slot=0):
        if str_mask:
            raise ValueError("mask files are not read here; pass the mask array via mask=")
        self._fe = _frontend or Frontend(im_height, im_width, n_levels, scale_factor, FAST_N_MIN, FAST_N_MAX,
                                         th_FAST_MIN, th_FAST_MAX, tile_h, tile_w, fixed_multi_scale_tile_size,
                                         apply_nms_ms, nms_ms_mode_gpu, mask=mask, device=device, max_images=1)
        self._slot = _slot
        self.n_levels_ = n_levels
        self.scale_factor_ = scale_factor

    def extract(self, image):
        self._fe.set_images(np.ascontiguousarray(image, np.uint8), first_slot=self._slot)
        self._fe.extract(self._slot, 1)
        return self._fe.get_keypoints(self._slot)

    def get_levels(self): return self.n_levels_
    def get_scale_factor(self): return self.scale_factor_
    def get_scale_factors(self): return self._fe.scale.copy()
    def get_inverse_scale_factors(self): return self._fe.inv_scale.copy()
    def get_scale_sigma_squares(self): return (self._fe.scale * self._fe.scale).astype(np.float32)
    def get_inverse_scale_sigma_squares(self): return (np.float32(1.0) / (self._fe.scale * self._fe.scale)).astype(np.float32)


class StereoORB:
    """Left/right extractor pair sharing one handle (slots 0/1) + the stereo matcher, i.e. what
    Frame::Frame(imLeft, imRight, ...) does on the hot path (src/Frame.cpp:80-250)."""

    def __init__(self, cfg, device=0):
        self.cfg = cfg
        self.fe = Frontend(**cfg.extractor_kwargs(), device=device, max_images=2)

    def __call__(self, im_left, im_right, th_high=100, th_low=50):
        self.fe.set_images(np.stack([im_left, im_right]).astype(np.uint8, copy=False))
        self.fe.extract(0, 2)
        self.fe.stereo_match(self.cfg.mb, self.cfg.mbf, 0, 1, th_high, th_low)
        kl, dl = self.fe.get_keypoints(0)
        kr, dr = self.fe.get_keypoints(1)
        ur, dp, bi, bd = self.fe.get_stereo(0)
        return {"kps_l": kl, "desc_l": dl, "kps_r": kr, "desc_r": dr, "u_right": ur, "depth": dp, "best_idx_r": bi,
                "best_dist": bd}


# ---- adjacent rows: stateless helpers on DEVICE arrays (torch CUDA tensors are the plumbing here) ----------
def _ptr(t):
    return C.c_void_p(t.data_ptr())


def project_points(P, Rcw, tcw, fx, fy, cx, cy, min_x, max_x, min_y, max_y, stream=None):
    """Mirror of orb_cuda::ORB_Search_by_projection_project_on_frame (include/cuda/orb_matcher.hpp:11-17).
    P: [3, n] f32 CUDA tensor (Px|Py|Pz planes), Rcw [9], tcw [3] f32 CUDA.  Returns (u, v, invz, is_valid) CUDA tensors."""
    import torch
    n = P.shape[1]
    u, v, iz = (torch.empty(n, dtype=torch.float32, device=P.device) for _ in range(3))
    ok = torch.empty(n, dtype=torch.uint8, device=P.device)
    _check(lib().jsfe_project_points(n, _ptr(P[0]), _ptr(P[1]), _ptr(P[2]), _ptr(Rcw), _ptr(tcw), fx, fy, cx, cy, min_x, max_x,
                                     min_y, max_y, _ptr(u), _ptr(v), _ptr(iz), _ptr(ok), _stream_ptr(stream)))
    return u, v, iz, ok


def hamming_pairs(idx_l, idx_r, desc_l, desc_r, stream=None):
    """Mirror of orb_cuda::ORB_compute_distances (include/cuda/orb_matcher.hpp:19-23): int32 CUDA index tensors, u8 [N,32] descriptors."""
    import torch
    n = idx_l.shape[0]
    d = torch.empty(n, dtype=torch.int32, device=idx_l.device)
    _check(lib().jsfe_hamming_pairs(n, _ptr(idx_l), _ptr(idx_r), _ptr(desc_l), _ptr(desc_r), _ptr(d), _stream_ptr(stream)))
    return d


def in_frustum(P, Pn, max_distance, inv_max, inv_min, Rcw, tcw, Ow, fx, fy, cx, cy, min_x, max_x, min_y, max_y, n_levels,
               log_scale_factor, view_cos_angle, stream=None):
    """Mirror of tracking_cuda::compute_isInFrustum_GPU (include/cuda/tracking_gpu.hpp:13-28).  Outputs other than the flag are
    defined only where the flag is 1 (as in the reference); they are zero-initialised here."""
    import torch
    n = P.shape[1]
    iz, u, v, vc = (torch.zeros(n, dtype=torch.float32, device=P.device) for _ in range(4))
    lvl = torch.zeros(n, dtype=torch.int32, device=P.device)
    ok = torch.empty(n, dtype=torch.uint8, device=P.device)
    _check(lib().jsfe_in_frustum(n, _ptr(P[0]), _ptr(P[1]), _ptr(P[2]), _ptr(Pn[0]), _ptr(Pn[1]), _ptr(Pn[2]), _ptr(max_distance),
                                 _ptr(inv_max), _ptr(inv_min), _ptr(Rcw), _ptr(tcw), _ptr(Ow), fx, fy, cx, cy, min_x, max_x, min_y,
                                 max_y, n_levels, log_scale_factor, view_cos_angle, _ptr(iz), _ptr(u), _ptr(v), _ptr(lvl), _ptr(vc),
                                 _ptr(ok), _stream_ptr(stream)))
    return iz, u, v, lvl, vc, ok


class MapPool:
    """Resident map-point SoA (jsfe_mappool_*): update() when map points change, in_frustum() per frame with a device id list."""

    def __init__(self, capacity, device=0):
        self._p = C.c_void_p()
        self.capacity, self.device = capacity, device
        _check(lib().jsfe_mappool_create(capacity, device, C.byref(self._p)))

    def update(self, ids, P, Pn, max_distance, inv_max, inv_min, stream=None):
        """Host arrays: ids int32 [n]; P, Pn float32 [3, n]; the three distance arrays float32 [n]."""
        a = lambda v, t: np.ascontiguousarray(v, t)
        ids, P, Pn = a(ids, np.int32), a(P, np.float32), a(Pn, np.float32)
        md, ima, imi = a(max_distance, np.float32), a(inv_max, np.float32), a(inv_min, np.float32)
        p = lambda x: C.c_void_p(x.ctypes.data)
        _check(lib().jsfe_mappool_update(self._p, len(ids), p(ids), p(P[0]), p(P[1]), p(P[2]), p(Pn[0]), p(Pn[1]), p(Pn[2]), p(md), p(ima), p(imi),
                                         _stream_ptr(stream)))

    def in_frustum(self, ids_dev, Rcw, tcw, Ow, fx, fy, cx, cy, min_x, max_x, min_y, max_y, n_levels, log_scale_factor, view_cos_angle, stream=None):
        """ids_dev: int32 CUDA tensor; Rcw [9], tcw [3], Ow [3]: host float32.  Returns the same tuple as in_frustum()."""
        import torch
        n = ids_dev.shape[0]
        dev = ids_dev.device
        iz, u, v, vc = (torch.zeros(n, dtype=torch.float32, device=dev) for _ in range(4))
        lvl = torch.zeros(n, dtype=torch.int32, device=dev)
        ok = torch.empty(n, dtype=torch.uint8, device=dev)
        a = lambda x: np.ascontiguousarray(x, np.float32)
        R, t, O = a(Rcw), a(tcw), a(Ow)
        p = lambda x: C.c_void_p(x.ctypes.data)
        _check(lib().jsfe_mappool_in_frustum(self._p, n, _ptr(ids_dev), p(R), p(t), p(O), fx, fy, cx, cy, min_x, max_x, min_y, max_y, n_levels,
                                             log_scale_factor, view_cos_angle, _ptr(iz), _ptr(u), _ptr(v), _ptr(lvl), _ptr(vc), _ptr(ok),
                                             _stream_ptr(stream)))
        return iz, u, v, lvl, vc, ok

    def close(self):
        if self._p:
            lib().jsfe_mappool_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------- SURVEY 8(f1)
class _SbpArgs(C.Structure):
    """ctypes image of jsfe_sbp_args (include/jsfe.h)."""
    _fields_ = ([("n_last", C.c_int32)] + [(n, C.c_void_p) for n in ("px", "py", "pz", "last_octave", "last_angle", "last_desc", "rcw9", "tcw3")] +
                [(n, C.c_float) for n in ("fx", "fy", "cx", "cy", "min_x", "max_x", "min_y", "max_y", "mbf", "th")] +
                [("scale_factors", C.c_float * 16), ("level_mode", C.c_int32), ("n_cur", C.c_int32)] +
                [(n, C.c_void_p) for n in ("cur_x", "cur_y", "cur_octave", "cur_angle", "cur_uright", "cur_occupied", "cur_desc",
                                           "cell_start", "cell_items")] +
                [("th_high", C.c_int32), ("check_orientation", C.c_int32)] +
                [(n, C.c_void_p) for n in ("best_idx2", "best_dist", "rot_bin", "cur_match", "hist", "n_matches")])


FRAME_GRID_COLS, FRAME_GRID_ROWS, HISTO_LENGTH = 64, 48, 30


def build_frame_grid(cur_x, cur_y, min_x, max_x, min_y, max_y, stream=None):
    """Device Frame::AssignFeaturesToGrid (src/Frame.cpp:464-479): -> (cell_start int32[64*48+1], cell_items int32[n]) CUDA tensors."""
    import torch
    n = cur_x.shape[0]
    start = torch.empty(FRAME_GRID_COLS * FRAME_GRID_ROWS + 1, dtype=torch.int32, device=cur_x.device)
    items = torch.full((max(n, 1),), -1, dtype=torch.int32, device=cur_x.device)
    _check(lib().jsfe_build_frame_grid(n, _ptr(cur_x), _ptr(cur_y), float(min_x), float(max_x), float(min_y), float(max_y),
                                       _ptr(start), _ptr(items), _stream_ptr(stream)))
    return start, items


def search_by_projection(last, cur, Rcw, tcw, fx, fy, cx, cy, min_x, max_x, min_y, max_y, mbf, th, scale_factors, level_mode,
                         th_high=100, check_orientation=True, grid=None, stream=None):
    """Device ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cpp:1647-1963) in one pass.
    last: dict of CUDA tensors P [3,n] f32, octave i32, angle f32 (deg), desc u8 [n,32];
    cur : dict of CUDA tensors x, y f32, octave i32, angle f32, uright f32, occupied u8 (or None), desc u8 [m,32];
    level_mode 0/1/2 = neither / bForward / bBackward.  grid: result of build_frame_grid (built here when None).
    -> dict of CUDA tensors best_idx2, best_dist, rot_bin [n], cur_match [m], hist [30], n_matches [1] (no synchronisation)."""
    import torch
    P = last["P"]
    dev = P.device
    n, m = P.shape[1], cur["x"].shape[0]
    if grid is None:
        grid = build_frame_grid(cur["x"], cur["y"], min_x, max_x, min_y, max_y, stream)
    out = dict(best_idx2=torch.empty(max(n, 1), dtype=torch.int32, device=dev), best_dist=torch.empty(max(n, 1), dtype=torch.int32, device=dev),
               rot_bin=torch.empty(max(n, 1), dtype=torch.int32, device=dev), cur_match=torch.empty(max(m, 1), dtype=torch.int32, device=dev),
               hist=torch.empty(HISTO_LENGTH, dtype=torch.int32, device=dev), n_matches=torch.empty(1, dtype=torch.int32, device=dev))
    a = _SbpArgs()
    a.n_last, a.n_cur = n, m
    a.px, a.py, a.pz = _ptr(P[0]), _ptr(P[1]), _ptr(P[2])
    a.last_octave, a.last_angle, a.last_desc = _ptr(last["octave"]), _ptr(last["angle"]), _ptr(last["desc"])
    a.rcw9, a.tcw3 = _ptr(Rcw), _ptr(tcw)
    a.fx, a.fy, a.cx, a.cy, a.min_x, a.max_x, a.min_y, a.max_y, a.mbf, a.th = fx, fy, cx, cy, min_x, max_x, min_y, max_y, mbf, th
    sf = np.zeros(16, np.float32)
    sf[:len(scale_factors)] = np.asarray(scale_factors, np.float32)
    a.scale_factors = (C.c_float * 16)(*sf.tolist())
    a.level_mode = int(level_mode)
    a.cur_x, a.cur_y, a.cur_octave, a.cur_angle = _ptr(cur["x"]), _ptr(cur["y"]), _ptr(cur["octave"]), _ptr(cur["angle"])
    a.cur_uright, a.cur_desc = _ptr(cur["uright"]), _ptr(cur["desc"])
    a.cur_occupied = _ptr(cur["occupied"]) if cur.get("occupied") is not None else None
    a.cell_start, a.cell_items = _ptr(grid[0]), _ptr(grid[1])
    a.th_high, a.check_orientation = int(th_high), int(bool(check_orientation))
    for k in ("best_idx2", "best_dist", "rot_bin", "cur_match", "hist", "n_matches"):
        setattr(a, k, _ptr(out[k]))
    _check(lib().jsfe_search_by_projection(C.byref(a), _stream_ptr(stream)))
    out["best_idx2"], out["best_dist"], out["rot_bin"], out["cur_match"] = out["best_idx2"][:n], out["best_dist"][:n], out["rot_bin"][:n], out["cur_match"][:m]
    out["grid"] = grid
    return out


# ---------------------------------------------------------------------------------------------------- SURVEY 8(f3)
def remap_bilinear(src, map_x, map_y, out=None, stream=None):
    """Device cv::remap(src, map_x, map_y, INTER_LINEAR) (Examples/Stereo/stereo_euroc.cpp:145-146), bit-exact with OpenCV 4.x.
    src: u8 CUDA tensor [n, h, w] (or [h, w]); map_x, map_y: contiguous f32 CUDA tensors [H, W] shared by the n images.
    out: optional u8 CUDA tensor [n, H, W'] view to write into (row/image strides are taken from it). -> out."""
    import torch
    one = src.dim() == 2
    s3 = src.unsqueeze(0) if one else src
    assert s3.stride(2) == 1 and map_x.is_contiguous() and map_y.is_contiguous() and map_x.shape == map_y.shape
    n, h, w = s3.shape
    H, W = map_x.shape
    if out is None:
        out = torch.empty((n, H, W), dtype=torch.uint8, device=src.device)
    o3 = out.unsqueeze(0) if out.dim() == 2 else out
    assert o3.stride(2) == 1 and o3.shape[0] == n and o3.shape[1] == H and o3.shape[2] >= W
    _check(lib().jsfe_remap_bilinear(_ptr(s3), h, w, s3.stride(1), s3.stride(0) if n > 1 else h * s3.stride(1), n, _ptr(map_x), _ptr(map_y),
                                     H, W, _ptr(o3), o3.stride(1), o3.stride(0) if n > 1 else H * o3.stride(1), _stream_ptr(stream)))
    return out[0] if (one and out.dim() == 3) else out


def cvt_gray(img, rgb=False, stream=None):
    """Device cv::cvtColor(img, *2GRAY) (src/Tracking.cpp:260-285): img u8 CUDA tensor [h, w, 3|4] contiguous -> [h, w]."""
    import torch
    assert img.is_contiguous() and img.shape[2] in (3, 4)
    h, w, c = img.shape
    out = torch.empty((h, w), dtype=torch.uint8, device=img.device)
    _check(lib().jsfe_cvt_gray(_ptr(img), h, w, w * c, c, int(bool(rgb)), _ptr(out), w, _stream_ptr(stream)))
    return out


# ---------------------------------------------------------------------------------------------------- SURVEY 8(f4)
CV_KEYPOINT_DTYPE = np.dtype([("x", np.float32), ("y", np.float32), ("size", np.float32), ("angle", np.float32),
                              ("response", np.float32), ("octave", np.int32), ("class_id", np.int32)])   # cv::KeyPoint, 28 bytes


def frame_view(fe, slot, stream=None):
    """Device Frame::Frame unpack (src/Frame.cpp:116-196) of slot `slot` of Frontend `fe`.
    -> dict of CUDA tensors: keys u8 [cap, 28] (cv::KeyPoint records), x, y, angle f32 [cap], octave i32 [cap]; valid: first n."""
    import torch
    dev = torch.device("cuda", fe.device)
    cap = fe.max_kp
    out = dict(keys=torch.zeros((cap, 28), dtype=torch.uint8, device=dev), x=torch.zeros(cap, dtype=torch.float32, device=dev),
               y=torch.zeros(cap, dtype=torch.float32, device=dev), octave=torch.zeros(cap, dtype=torch.int32, device=dev),
               angle=torch.zeros(cap, dtype=torch.float32, device=dev))
    _check(lib().jsfe_frame_view(fe._h, slot, _ptr(out["keys"]), _ptr(out["x"]), _ptr(out["y"]), _ptr(out["octave"]), _ptr(out["angle"]),
                                 _stream_ptr(stream)))
    return out


class DevicePtr:
    """A raw device pointer owned by a handle (e.g. a jsfe_slot_view field), usable wherever the bindings take a CUDA tensor."""

    def __init__(self, ptr):
        self._p = int(ptr or 0)

    def data_ptr(self):
        return self._p
