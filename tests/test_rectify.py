"""the input side -- cv::remap(INTER_LINEAR) rectification and cv::cvtColor(*2GRAY) on the device.

The algorithm is OpenCV's (third-party to the reference, which links libopencv 4.1; called at
Examples/Stereo/stereo_euroc.cpp:106-107,145-146 and src/Tracking.cpp:260-285).
CPU part : the C oracle against golden vectors written from cv2 (tools/make_golden_cv.py) and, when the cv2 wheel is
           importable, against cv2 live on fresh random inputs.
GPU part : the CUDA kernels through the C ABI, bit-exact against the oracle, the golden vectors and (live) cv2, including
           the batch form that rectifies straight into the extractor's level-0 slots followed by a full extraction."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle import oracle as orc

GOLD = os.path.join(ROOT, "tests", "golden")


def _wild_case(seed, hs=97, ws=131, hd=80, wd=123):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, 256, size=(hs, ws), dtype=np.uint8)
    mx = rng.uniform(-10, ws + 10, size=(hd, wd)).astype(np.float32)
    my = rng.uniform(-10, hs + 10, size=(hd, wd)).astype(np.float32)
    mx[::3] = np.round(mx[::3]); my[::4] = np.round(my[::4]); mx[::5] += np.float32(0.5)
    mx[2, :4] = [ws - 1, ws - 1.5, -1, -0.5]
    my[2, :4] = [hs - 1, hs - 1, 0, 0]
    return src, mx, my


def test_oracle_remap_matches_golden_vectors():
    g = np.load(os.path.join(GOLD, "cv_remap.npz"))
    for k in "ab":
        got = orc.remap_bilinear(g[k + "_src"], g[k + "_mx"], g[k + "_my"])
        assert np.array_equal(got, g[k + "_dst"]), k


def test_oracle_gray_matches_golden_vectors():
    g = np.load(os.path.join(GOLD, "cv_gray.npz"))
    img = g["img"]
    assert np.array_equal(orc.cvt_gray(img[..., :3], rgb=False), g["bgr"])
    assert np.array_equal(orc.cvt_gray(img[..., :3], rgb=True), g["rgb"])
    assert np.array_equal(orc.cvt_gray(img, rgb=False), g["bgra"])
    assert np.array_equal(orc.cvt_gray(img, rgb=True), g["rgba"])


def test_oracle_remap_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    for seed in range(4):
        src, mx, my = _wild_case(seed)
        assert np.array_equal(orc.remap_bilinear(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR)), seed
    # identity map reproduces the image; a constant shift by a whole pixel is a plain copy with a zero border
    src = np.random.default_rng(9).integers(0, 256, size=(40, 50), dtype=np.uint8)
    xs, ys = np.meshgrid(np.arange(50, dtype=np.float32), np.arange(40, dtype=np.float32))
    assert np.array_equal(orc.remap_bilinear(src, xs, ys), src)
    sh = orc.remap_bilinear(src, xs + 3, ys)
    assert np.array_equal(sh[:, :47], src[:, 3:]) and (sh[:, 47:] == 0).all()


# ---------------------------------------------------------------------------------------------------------- GPU tests
@pytest.mark.gpu
def test_cuda_remap_and_gray_match_oracle_and_golden():
    import torch
    from jetson_slam_b200 import frontend
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g = np.load(os.path.join(GOLD, "cv_remap.npz"))
    for k in "ab":
        t = (d(g[k + "_src"]), d(g[k + "_mx"]), d(g[k + "_my"]))
        out = frontend.remap_bilinear(*t)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), g[k + "_dst"]), k
    for seed in range(3):
        src, mx, my = _wild_case(10 + seed, hs=201, ws=333, hd=190, wd=301)     # odd width: byte-store tail path
        t = (d(src), d(mx), d(my))
        out = frontend.remap_bilinear(*t)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), orc.remap_bilinear(src, mx, my)), seed
    gg = np.load(os.path.join(GOLD, "cv_gray.npz"))
    img = gg["img"]
    for ch, rgb, key in ((3, False, "bgr"), (3, True, "rgb"), (4, False, "bgra"), (4, True, "rgba")):
        t = d(img[..., :ch])
        out = frontend.cvt_gray(t, rgb=rgb)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), gg[key]), key


@pytest.mark.gpu
def test_cuda_rectify_into_slots_then_extract():
    """Raw frames -> device rectification straight into the level-0 slots -> extraction == oracle on cv2/oracle-rectified frames."""
    import ctypes as C
    import torch
    from jetson_slam_b200 import frontend, synth
    from jetson_slam_b200.configs import CONFIGS
    cfg = CONFIGS["C1"]                                     # 320x240
    h, w = cfg.height, cfg.width
    rng = np.random.default_rng(3)
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    r2 = ((xs - w / 2) ** 2 + (ys - h / 2) ** 2) / np.float32(w * w)
    mx = (xs + (xs - w / 2) * 0.08 * r2 + 1.3).astype(np.float32)          # mild radial distortion + shift
    my = (ys + (ys - h / 2) * 0.08 * r2 - 0.7).astype(np.float32)
    raw = np.stack([synth.stereo_pair(h, w, 40 + i)[0] for i in range(3)])  # three frames of one camera
    want_rect = np.stack([orc.remap_bilinear(f, mx, my) for f in raw])
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=6)
    lib = frontend.lib()
    p0, p2, pitch = C.c_void_p(), C.c_void_p(), C.c_int64()
    assert lib.jsfe_slot_image(fe._h, 0, C.byref(p0), C.byref(pitch)) == 0
    assert lib.jsfe_slot_image(fe._h, 2, C.byref(p2), None) == 0
    stride2 = p2.value - p0.value                            # every second slot (= the left eyes)
    traw, tmx, tmy = torch.from_numpy(raw).cuda(), torch.from_numpy(mx).cuda(), torch.from_numpy(my).cuda()
    rc = lib.jsfe_remap_bilinear(C.c_void_p(traw.data_ptr()), h, w, w, h * w, 3, C.c_void_p(tmx.data_ptr()), C.c_void_p(tmy.data_ptr()),
                                 h, w, p0, pitch.value, stride2, None)
    assert rc == 0, lib.jsfe_last_error()
    torch.cuda.synchronize()
    fe.extract(0, 6)
    o = orc.Oracle(**cfg.extractor_kwargs())
    for i in range(3):
        assert np.array_equal(fe.level_image(2 * i, 0), want_rect[i]), i
        kps, desc = fe.get_keypoints(2 * i)
        wk, wd = o.extract(want_rect[i])
        assert np.array_equal(kps, wk) and np.array_equal(desc, wd), i
    fe.close()


def test_oracle_remap_property_against_cv2():
    """Hypothesis sweep over small shapes and map values (in range, on the border, far outside, exact grid points)."""
    cv2 = pytest.importorskip("cv2")
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 23), st.integers(1, 29), st.integers(1, 17), st.integers(1, 19), st.integers(0, 2 ** 31 - 1))
    def check(hs, ws, hd, wd, seed):
        rng = np.random.default_rng(seed)
        src = rng.integers(0, 256, size=(hs, ws), dtype=np.uint8)
        mx = rng.uniform(-3, ws + 3, size=(hd, wd)).astype(np.float32)
        my = rng.uniform(-3, hs + 3, size=(hd, wd)).astype(np.float32)
        snap = rng.random((hd, wd)) < 0.3
        mx[snap] = np.round(mx[snap] * 2) / 2
        my[snap] = np.round(my[snap])
        assert np.array_equal(orc.remap_bilinear(src, mx, my), cv2.remap(src, mx, my, cv2.INTER_LINEAR))

    check()
