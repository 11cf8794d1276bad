"""CPU tests of the oracle (test infrastructure): pinned against golden vectors produced by the REFERENCE's own
src/cuda compiled unmodified for sm_100a and run on a B200 (tools/make_golden.py, tests/golden/ref_*.npz)."""
import glob
import os

import numpy as np
import pytest

from jetson_slam_b200 import synth
from jetson_slam_b200.configs import CONFIGS
from oracle import oracle as orc
from conftest import GOLDEN


def _golden_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_*_seed*.npz"))):
        name, seed = os.path.basename(f)[4:-4].rsplit("_seed", 1)
        out.append((name, int(seed), f))
    return out


@pytest.mark.parametrize("name,seed,path", _golden_cases(), ids=lambda v: str(v) if not isinstance(v, str) or len(v) < 20 else "")
def test_oracle_matches_reference_golden(name, seed, path):
    cfg = CONFIGS[name]
    g = np.load(path)
    L, R = synth.stereo_pair(cfg.height, cfg.width, seed)
    import hashlib
    assert hashlib.sha256(L.tobytes()).hexdigest() == str(g["img_l_sha"]), "synthetic input drifted"
    ol, orr = orc.Oracle(**cfg.extractor_kwargs()), orc.Oracle(**cfg.extractor_kwargs())
    kl, dl = ol.extract(L)
    kr, dr = orr.extract(R)
    # bit-exact: coordinates, scores, angle bits, octave, size, 256-bit descriptors
    assert kl.shape == g["kps_l"].shape and np.array_equal(kl, g["kps_l"])
    assert kr.shape == g["kps_r"].shape and np.array_equal(kr, g["kps_r"])
    assert np.array_equal(dl, g["desc_l"]) and np.array_equal(dr, g["desc_r"])
    assert np.array_equal(ol.n_keypoints(), g["n_per_level_l"])
    # intermediates (digests of the reference's level images / blurred images / score maps)
    for l in range(cfg.n_levels):
        assert hashlib.sha256(ol.level_image(l).tobytes()).hexdigest() == g["img_sha_l"][l]
        assert hashlib.sha256(ol.level_blur(l).tobytes()).hexdigest() == g["blur_sha_l"][l]
        assert hashlib.sha256(ol.level_score(l).tobytes()).hexdigest() == g["score_sha_l"][l]
    ur, dp, bi, bd = orc.stereo_match(ol, orr, kl, dl, kr, dr, cfg.mb, cfg.mbf)
    # north_star tolerance for sub-pixel disparity is 1e-4; the oracle is in fact bit-exact
    assert np.array_equal(ur.view(np.int32), g["u_right"].view(np.int32))
    assert np.array_equal(dp.view(np.int32), g["depth"].view(np.int32))
    assert (ur >= 0).sum() > 0


def test_oracle_degenerate_golden():
    cfg = CONFIGS["C1"]
    g = np.load(os.path.join(GOLDEN, "ref_degenerate_C1.npz"))
    o = orc.Oracle(**cfg.extractor_kwargs())
    for nm, img in synth.degenerate_images(cfg.height, cfg.width).items():
        k, d = o.extract(img)
        assert k.shape == g[f"kps_{nm}"].shape and np.array_equal(k, g[f"kps_{nm}"]), nm
        assert np.array_equal(d, g[f"desc_{nm}"]), nm
    assert g["kps_zeros"].shape[1] == 0 and g["kps_checker"].shape[1] > 0


def test_lut_matches_bruteforce_spec():
    """A.3: the scan procedure is the spec; it differs from 'maximal circular run in band' on exactly 6 masks for (9,14)."""
    lut = orc.Oracle(64, 64, n_levels=1, tile_h=16, tile_w=16).lut()
    assert lut.sum() == 1014 and lut[0xFFFF] == 0

    def circular_runs(m):
        if m == 0xFFFF:
            return [16]
        bits = [(m >> (15 - i)) & 1 for i in range(16)]
        s = bits.index(0)
        bits = bits[s:] + bits[:s]
        runs, r = [], 0
        for b in bits:
            if b:
                r += 1
            else:
                if r:
                    runs.append(r)
                r = 0
        if r:
            runs.append(r)
        return runs
    diff = [m for m in range(0xFFFF) if bool(lut[m]) != any(9 <= r <= 14 for r in circular_runs(m))]
    assert sorted(diff) == [0xFFBF, 0xFFDF, 0xFFEF, 0xFFF7, 0xFFFB, 0xFFFD]


@pytest.mark.parametrize("tw", list(range(1, 129)))
def test_column_rank_is_a_total_order(tw):
    """A.4: the smem tree's tie-break equals a fixed column priority; check on random multi-way ties."""
    rank = orc.column_rank(tw)
    assert sorted(rank.tolist()) == list(range(tw))
    rng = np.random.default_rng(tw)
    for _ in range(40):
        val = rng.integers(0, 3, size=tw)
        # literal tree
        v, ident = val.copy(), np.arange(tw)
        g = (tw - 1) // 2 + 1
        for _it in range(int(np.ceil(np.log2(np.float32(tw)))) if tw > 1 else 0):
            for j in range(min(g, tw)):
                if j + g < tw and v[j] < v[j + g]:
                    v[j], ident[j] = v[j + g], ident[j + g]
            g = (g - 1) // 2 + 1
        m = val.max()
        expect = min((j for j in range(tw) if val[j] == m), key=lambda j: rank[j])
        assert ident[0] == expect


def test_libdevice_transcriptions_close_to_libm():
    rng = np.random.default_rng(0)
    L = orc.lib()
    for _ in range(2000):
        a = np.float32(rng.uniform(-np.pi, np.pi))
        assert abs(L.orc_cosf(a) - np.cos(np.float64(a))) < 3e-7
        assert abs(L.orc_sinf(a) - np.sin(np.float64(a))) < 3e-7
        y, x = np.float32(rng.integers(-3e6, 3e6)), np.float32(rng.integers(-3e6, 3e6))
        assert abs(L.orc_atan2f(y, x) - np.arctan2(np.float64(y), np.float64(x))) < 6e-7
    assert L.orc_atan2f(0.0, 0.0) == 0.0 and abs(L.orc_atan2f(0.0, -1.0) - np.pi) < 1e-6


def test_geometry_tables_match_survey():
    o = orc.Oracle(**CONFIGS["C2"].extractor_kwargs())
    assert o.max_kp == 2016
    assert list(zip(o.w, o.h)) == [(1241, 376), (1034, 313), (861, 261), (718, 217), (598, 181), (498, 151), (415, 125), (346, 104)]
    assert o.umax().tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert CONFIGS["C1"].height == 240 and orc.Oracle(**CONFIGS["C1"].extractor_kwargs()).max_kp == 504
    assert orc.Oracle(**CONFIGS["C5"].extractor_kwargs()).max_kp == 4029


def test_mask_suppresses_keypoints():
    cfg = CONFIGS["tiny"]
    img, _ = synth.stereo_pair(cfg.height, cfg.width, 5)
    mask = np.full((cfg.height, cfg.width), 255, np.uint8)
    mask[:, : cfg.width // 2] = 0
    k, _ = orc.Oracle(**cfg.extractor_kwargs(), mask=mask).extract(img)
    k0, _ = orc.Oracle(**cfg.extractor_kwargs()).extract(img)
    assert k.shape[1] > 0 and k.shape[1] < k0.shape[1]
    assert (k[0][k[4] == 0] >= cfg.width // 2 - 1).all()


def test_nms_ms_modes_run_and_only_remove():
    base = CONFIGS["tiny"].extractor_kwargs()
    img, _ = synth.stereo_pair(120, 160, 9)
    n0 = orc.Oracle(**base).extract(img)[0].shape[1]
    for mode in (0, 1):
        kw = dict(base, apply_nms_ms=1, nms_ms_mode_gpu=mode)
        n = orc.Oracle(**kw).extract(img)[0].shape[1]
        assert 0 < n <= n0
