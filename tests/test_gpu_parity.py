"""GPU parity tests (run on the B200 with -m gpu): the CUDA path, called through the C ABI, against
 (1) the CPU oracle on the same seeded inputs, stage by stage, and
 (2) the committed golden vectors produced by the reference's own kernels (tests/golden/ref_*.npz).
Bar: bit-exact for coordinates, scores, angles, descriptors, Hamming arg-min; u_right/depth are asserted
bit-exact too (north_star allows 1e-4)."""
import glob
import os

import numpy as np
import pytest

from conftest import GOLDEN
from jetson_slam_b200 import frontend, synth
from jetson_slam_b200.configs import CONFIGS
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _oracle_pair(cfg, L, R):
    ol, orr = orc.Oracle(**cfg.extractor_kwargs()), orc.Oracle(**cfg.extractor_kwargs())
    kl, dl = ol.extract(L)
    kr, dr = orr.extract(R)
    st = orc.stereo_match(ol, orr, kl, dl, kr, dr, cfg.mb, cfg.mbf)
    return ol, orr, (kl, dl, kr, dr), st


def _assert_pair_equal(out, kl, dl, kr, dr, st, tag=""):
    ur, dp, bi, bd = st
    assert out["kps_l"].shape == kl.shape, f"{tag} N_left {out['kps_l'].shape} vs {kl.shape}"
    assert out["kps_r"].shape == kr.shape, f"{tag} N_right"
    for r, nm in enumerate(("x", "y", "score", "angle", "octave", "size")):
        assert np.array_equal(out["kps_l"][r], kl[r]), f"{tag} left {nm}"
        assert np.array_equal(out["kps_r"][r], kr[r]), f"{tag} right {nm}"
    assert np.array_equal(out["desc_l"], dl), f"{tag} left descriptors: {np.unpackbits(out['desc_l'] ^ dl).sum()} bits differ"
    assert np.array_equal(out["desc_r"], dr), f"{tag} right descriptors"
    if bi is not None:
        assert np.array_equal(out["best_idx_r"], bi), f"{tag} Hamming arg-min"
        assert np.array_equal(out["best_dist"], bd), f"{tag} Hamming distance"
    assert np.array_equal(out["u_right"] >= 0, ur >= 0), f"{tag} match mask"
    assert np.abs(out["u_right"] - ur).max(initial=0) <= 1e-4 and np.abs(out["depth"] - dp).max(initial=0) <= 1e-4 * max(1.0, np.abs(dp).max(initial=0))
    assert np.array_equal(out["u_right"].view(np.int32), ur.view(np.int32)), f"{tag} u_right bits"
    assert np.array_equal(out["depth"].view(np.int32), dp.view(np.int32)), f"{tag} depth bits"


@pytest.mark.parametrize("name,seed", [("tiny", 0), ("tiny", 1), ("tiny-fixed", 0), ("C1", 0), ("C2", 0), ("C3", 0), ("C4", 0),
                                       ("KITTI00-02", 0), ("EuRoC", 0), ("C5", 0),
                                       ("KITTI04-12", 0), ("KAIST-nmsms-cpu", 0)])
def test_stages_match_oracle(name, seed):
    cfg = CONFIGS[name]
    L, R = synth.stereo_pair(cfg.height, cfg.width, seed)
    ol, orr, (kl, dl, kr, dr), st = _oracle_pair(cfg, L, R)
    s = frontend.StereoORB(cfg)
    out = s(L, R)
    fe = s.fe
    for slot, o in ((0, ol), (1, orr)):
        for l in range(cfg.n_levels):
            assert np.array_equal(fe.level_image(slot, l), o.level_image(l)), f"pyramid level {l} slot {slot}"
            gb, ob = fe.level_blur(slot, l), o.level_blur(l)
            assert np.array_equal(gb, ob), f"blurred level {l} slot {slot}: {np.count_nonzero(gb != ob)} px differ"
        cx, cy, cs = fe.cells(slot)  # after the optional cross-scale NMS, as in the oracle
        ox, oy, os_ = o.cells()
        assert np.array_equal(cs, os_), f"cell scores slot {slot}: {np.count_nonzero(cs != os_)} differ"
        pos = os_ > 0
        assert np.array_equal(cx[pos], ox[pos]) and np.array_equal(cy[pos], oy[pos]), f"cell arg-max (tie-break) slot {slot}"
        x, y, sc, lv, ang = fe.level_keypoints(slot)
        n = o.n_keypoints()
        idx = np.concatenate([o.level_offset[l] + np.arange(n[l]) for l in range(cfg.n_levels)]).astype(np.int64)
        okx, oky, oks, oka = o.level_keypoints()
        N = len(idx)
        assert np.array_equal(x[:N], okx[idx]) and np.array_equal(y[:N], oky[idx]) and np.array_equal(sc[:N], oks[idx])
        assert np.array_equal(ang[:N].view(np.int32), oka[idx].view(np.int32)), "orientation bits"
    _assert_pair_equal(out, kl, dl, kr, dr, st, tag=name)
    assert (st[0] >= 0).sum() > 0


@pytest.mark.parametrize("n_min,n_max,th", [(12, 16, 20), (5, 8, 20), (3, 16, 10), (1, 4, 30), (16, 16, 20), (9, 9, 5)])
def test_other_arc_bands_match_oracle(n_min, n_max, th):
    """FAST_N_MIN decides which compass pre-test k_fast_cells may use (3 / 2 / 1 / 0 adjacent compass points); every band must
    stay exact, including the ones whose pre-test falls back to the reference's two early-outs only."""
    import dataclasses
    cfg = dataclasses.replace(CONFIGS["C1"], fast_n_min=n_min, fast_n_max=n_max, th_fast_max=th)
    L, R = synth.stereo_pair(cfg.height, cfg.width, 5)
    o = orc.Oracle(**cfg.extractor_kwargs())
    wk, wd = o.extract(L)
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=1)
    fe.set_images(L[None])
    fe.extract(0, 1)
    kps, desc = fe.get_keypoints(0)
    cx, cy, cs = fe.cells(0)
    ox, oy, os_ = o.cells()
    assert np.array_equal(cs, os_), f"cell scores: {np.count_nonzero(cs != os_)} differ"
    assert np.array_equal(kps, wk) and np.array_equal(desc, wd)
    if (n_min, n_max) == (16, 16):
        assert kps.shape[1] == 0          # the LUT never accepts 0xFFFF (orb_gpu.cpp:366-436)
    fe.close()


@pytest.mark.parametrize("cap", [1, 40, 700])
@pytest.mark.parametrize("kind", ["synthetic", "noise", "checker"])
def test_fast_list_overflow_paths_stay_exact(cap, kind, monkeypatch):
    """k_fast_cells keeps its survivors in a shared-memory work list sized by jsfe_create (LevelGeom::fast_cap).  JSFE_DEBUG_FAST_CAP
    shrinks it so that the overflow path runs on most or all tiles (dense evaluation of the tile, both polarities, then the dense NMS
    walk) and, at cap 700, beside tiles that fit (positives written over consumed entries); results must not change.
    Noise and checkerboard images also put survivors of BOTH polarities next to each other."""
    import dataclasses
    cfg = dataclasses.replace(CONFIGS["C1"], th_fast_max=12)
    if kind == "synthetic":
        img = synth.stereo_pair(cfg.height, cfg.width, 9)[0]
    elif kind == "noise":
        img = np.random.default_rng(3).integers(0, 256, size=(cfg.height, cfg.width), dtype=np.uint8)
    else:
        yy, xx = np.mgrid[0:cfg.height, 0:cfg.width]
        img = (((yy // 3 + xx // 3) % 2) * 200 + 20 + (yy * 7 + xx * 13) % 5).astype(np.uint8)
    o = orc.Oracle(**cfg.extractor_kwargs())
    wk, wd = o.extract(img)
    ox, oy, os_ = o.cells()
    monkeypatch.setenv("JSFE_DEBUG_FAST_CAP", str(cap))
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=1)
    monkeypatch.delenv("JSFE_DEBUG_FAST_CAP")
    fe.set_images(img[None])
    fe.extract(0, 1)
    cx, cy, cs = fe.cells(0)
    assert np.array_equal(cs, os_), f"cell scores: {np.count_nonzero(cs != os_)} differ"
    pos = os_ > 0
    assert np.array_equal(cx[pos], ox[pos]) and np.array_equal(cy[pos], oy[pos])
    kps, desc = fe.get_keypoints(0)
    assert np.array_equal(kps, wk) and np.array_equal(desc, wd)
    fe.close()


def _random_geometry(i):
    rng = np.random.default_rng(1000 + i)
    h, w = int(rng.integers(70, 420)), int(rng.integers(70, 520))
    levels = int(rng.integers(1, 9))
    scale = float(np.float32(rng.choice([1.1, 1.2, 1.3, 1.5, 2.0])))
    while levels > 1 and min(h, w) / scale ** (levels - 1) < 45:      # every level needs an interior
        levels -= 1
    tile = int(rng.integers(6, 90))
    fixed = int(rng.integers(0, 2))
    while not fixed and int(np.float32(tile) * np.float32(1.0 / scale ** (levels - 1))) < 1:
        tile += 4
    return dict(height=h, width=w, n_levels=levels, scale_factor=scale, tile_h=tile, tile_w=min(tile, 128),
                fixed_multi_scale_tile_size=fixed, apply_nms_ms=int(rng.integers(0, 2)), nms_ms_mode_gpu=int(rng.integers(0, 2)))


@pytest.mark.parametrize("i", range(16))
def test_random_geometries_match_oracle(i):
    """Odd sizes, tile sizes from 6 to 89 px, 1-8 levels, several scale factors, both cross-scale NMS rules: the block geometry
    of every kernel (TMA boxes, work maps, phase-A thread grid, tile-row index) is derived from these."""
    import dataclasses
    g = _random_geometry(i)
    cfg = dataclasses.replace(CONFIGS["tiny"], name=f"rand{i}", **g)
    L, R = synth.stereo_pair(cfg.height, cfg.width, 300 + i)
    ol, orr, (kl, dl, kr, dr), st = _oracle_pair(cfg, L, R)
    out = frontend.StereoORB(cfg)(L, R)
    _assert_pair_equal(out, kl, dl, kr, dr, st, tag=str(g))


def _golden_cases():
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_*_seed*.npz"))):
        name, seed = os.path.basename(f)[4:-4].rsplit("_seed", 1)
        out.append((name, int(seed)))
    return out


@pytest.mark.parametrize("name,seed", _golden_cases())
def test_matches_reference_golden(name, seed):
    """Same inputs as the reference run that produced the fixture -> identical outputs."""
    cfg = CONFIGS[name]
    g = np.load(os.path.join(GOLDEN, f"ref_{name}_seed{seed}.npz"))
    L, R = synth.stereo_pair(cfg.height, cfg.width, seed)
    out = frontend.StereoORB(cfg)(L, R)
    _assert_pair_equal(out, g["kps_l"], g["desc_l"], g["kps_r"], g["desc_r"], (g["u_right"], g["depth"], None, None), tag=name)


def test_degenerate_inputs_match_reference_golden():
    cfg = CONFIGS["C1"]
    g = np.load(os.path.join(GOLDEN, "ref_degenerate_C1.npz"))
    ex = frontend.ORBExtractor(cfg.height, cfg.width, cfg.scale_factor, cfg.n_levels, cfg.fast_n_min, cfg.fast_n_max,
                               cfg.th_fast_min, cfg.th_fast_max, "", cfg.tile_h, cfg.tile_w, False, False, True)
    for nm, img in synth.degenerate_images(cfg.height, cfg.width).items():
        k, d = ex.extract(img)
        assert k.shape == g[f"kps_{nm}"].shape and np.array_equal(k, g[f"kps_{nm}"]), nm
        assert np.array_equal(d, g[f"desc_{nm}"]), nm


def test_blur_is_exact_on_flat_and_saturated_images():
    """Constant / two-level / saturated windows put the separable blur value within 1e-5 of an integer, i.e. they all
    take the exact-chain fallback (the list pass at the end of k_blur); the stored bytes must still be the reference's."""
    cfg = CONFIGS["C1"]
    rng = np.random.default_rng(3)
    imgs = dict(synth.degenerate_images(cfg.height, cfg.width))
    imgs["flat200"] = np.full((cfg.height, cfg.width), 200, np.uint8)
    imgs["blocks"] = (rng.integers(0, 256, size=(cfg.height // 8 + 1, cfg.width // 8 + 1)).astype(np.uint8)
                      .repeat(8, 0).repeat(8, 1)[:cfg.height, :cfg.width]).copy()
    imgs["ramp"] = np.tile((np.arange(cfg.width) // 3 % 256).astype(np.uint8), (cfg.height, 1))
    o = orc.Oracle(**cfg.extractor_kwargs())
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=1)
    for nm, img in imgs.items():
        o.extract(img)
        fe.set_images(np.ascontiguousarray(img))
        fe.extract(0, 1)
        for l in range(cfg.n_levels):
            gb, ob = fe.level_blur(0, l), o.level_blur(l)
            assert np.array_equal(gb, ob), f"{nm} level {l}: {np.count_nonzero(gb != ob)} blurred px differ"
        k0, d0 = o.extract(img)
        k1, d1 = fe.get_keypoints(0)
        assert np.array_equal(k0, k1) and np.array_equal(d0, d1), nm


def test_empty_pair_gives_no_matches():
    cfg = CONFIGS["tiny"]
    z = np.zeros((cfg.height, cfg.width), np.uint8)
    out = frontend.StereoORB(cfg)(z, z)
    assert out["kps_l"].shape == (6, 0) and out["u_right"].shape == (0,)


def test_batch_slots_are_independent_and_order_invariant():
    """8 pairs in one launch == the same pairs one at a time (slots do not interact)."""
    cfg = CONFIGS["C1"]
    pairs = [synth.stereo_pair(cfg.height, cfg.width, s) for s in range(8)]
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=16)
    fe.set_images(np.stack([im for p in pairs for im in p]))
    fe.extract(0, 16)
    fe.stereo_match(cfg.mb, cfg.mbf, 0, 8)
    single = frontend.StereoORB(cfg)
    for p, (L, R) in enumerate(pairs):
        ref = single(L, R)
        kl, dl = fe.get_keypoints(2 * p)
        kr, dr = fe.get_keypoints(2 * p + 1)
        ur, dp, bi, bd = fe.get_stereo(p)
        assert np.array_equal(kl, ref["kps_l"]) and np.array_equal(kr, ref["kps_r"])
        assert np.array_equal(dl, ref["desc_l"]) and np.array_equal(dr, ref["desc_r"])
        assert np.array_equal(ur.view(np.int32), ref["u_right"].view(np.int32)) and np.array_equal(bi, ref["best_idx_r"])
    d = fe.download(0, 16)
    assert d["n"][0] == fe.get_keypoints(0)[0].shape[1] and d["bytes"] > 0
    # the pipelined end-to-end call (host images in, host slabs out) gives the same bytes as the staged calls
    want = {k: np.array(v) for k, v in d.items() if k != "bytes"}
    host = np.stack([im for p in pairs for im in p])
    for chunk in (3, 8):
        e = fe.process_host_pairs(host, cfg.mb, cfg.mbf, chunk_pairs=chunk)
        assert np.array_equal(e["n"], want["n"])
        for s_ in range(16):
            n = want["n"][s_]
            assert np.array_equal(e["kps"][s_, :, :n], want["kps"][s_, :, :n]) and np.array_equal(e["desc"][s_, :n], want["desc"][s_, :n])
            if s_ % 2 == 0:
                assert np.array_equal(e["u_right"][s_, :n].view(np.int32), want["u_right"][s_, :n].view(np.int32))
                assert np.array_equal(e["depth"][s_, :n].view(np.int32), want["depth"][s_, :n].view(np.int32))


@pytest.mark.parametrize("write_combined", [False, True])
def test_host_buffer_from_the_abi_allocator_feeds_the_host_call(write_combined):
    """jsfe_host_alloc (cached and write-combined pinned memory) as the `images` buffer of jsfe_process_host_pairs gives the same
    results as a numpy array, and jsfe_host_free releases it."""
    cfg = CONFIGS["C1"]
    pairs = [synth.stereo_pair(cfg.height, cfg.width, 40 + i) for i in range(3)]
    host = np.stack([im for p in pairs for im in p])
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=6)
    want = fe.process_host_pairs(host, cfg.mb, cfg.mbf)
    want = {k: np.array(v) for k, v in want.items() if k != "bytes"}
    hb = frontend.HostBuffer(host.shape, np.uint8, write_combined=write_combined)
    hb.array[...] = host
    got = fe.process_host_pairs(hb.array, cfg.mb, cfg.mbf)
    assert np.array_equal(got["n"], want["n"]) and want["n"].min() > 0
    for s_ in range(6):
        n = want["n"][s_]
        assert np.array_equal(got["kps"][s_, :, :n], want["kps"][s_, :, :n]) and np.array_equal(got["desc"][s_, :n], want["desc"][s_, :n])
        if s_ % 2 == 0:
            assert np.array_equal(got["u_right"][s_, :n].view(np.int32), want["u_right"][s_, :n].view(np.int32))
    hb.close()
    hb.close()      # idempotent
    fe.close()


def test_chunk_schedule_covers_every_pair_exactly_once():
    """jsfe_process_host_pairs for every (chunk_pairs, n_pairs) in 1..4 x 1..16: the ramped schedule (C/4, C/2, C ... C/2, C/4) must
    process exactly the caller's pairs (a schedule that ran past n_pairs would read beyond the host buffer) and give the bytes of
    the unchunked call."""
    cfg = CONFIGS["tiny"]
    pairs = [synth.stereo_pair(cfg.height, cfg.width, 40 + s) for s in range(16)]
    host = np.stack([im for p in pairs for im in p])
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=32)
    e = fe.process_host_pairs(host, cfg.mb, cfg.mbf, chunk_pairs=16)
    want = {k: np.array(v) for k, v in e.items() if k != "bytes"}
    for chunk in (1, 2, 3, 4):
        for n in range(1, 17):
            # the guard page: the buffer handed in ends right after pair n-1
            e = fe.process_host_pairs(np.ascontiguousarray(host[: 2 * n]), cfg.mb, cfg.mbf, chunk_pairs=chunk)
            assert np.array_equal(e["n"], want["n"][: 2 * n]), (chunk, n)
            for s_ in range(2 * n):
                k = want["n"][s_]
                assert np.array_equal(e["kps"][s_, :, :k], want["kps"][s_, :, :k]) and np.array_equal(e["desc"][s_, :k], want["desc"][s_, :k]), (chunk, n, s_)
                if s_ % 2 == 0:
                    assert np.array_equal(e["u_right"][s_, :k].view(np.int32), want["u_right"][s_, :k].view(np.int32)), (chunk, n, s_)
    fe.close()


def test_single_chunk_graph_replay_tracks_inputs_and_parameters(monkeypatch):
    """jsfe_process_host_pairs replays a CUDA graph when the call is one chunk: new images, a changed pair count and changed
    matcher parameters must all take effect (the graph bakes kernel arguments in), and it must equal the stream path."""
    cfg = CONFIGS["C1"]
    pairs = [synth.stereo_pair(cfg.height, cfg.width, 60 + s) for s in range(3)]
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=4)

    def run(imgs, mb, mbf, **kw):
        e = fe.process_host_pairs(np.stack(imgs), mb, mbf, **kw)
        return {k: np.array(v) for k, v in e.items() if k != "bytes"}

    def same(a, b):
        assert np.array_equal(a["n"], b["n"])
        for s_ in range(len(a["n"])):
            n = a["n"][s_]
            assert np.array_equal(a["kps"][s_, :, :n], b["kps"][s_, :, :n]) and np.array_equal(a["desc"][s_, :n], b["desc"][s_, :n])
            if s_ % 2 == 0:
                assert np.array_equal(a["u_right"][s_, :n].view(np.int32), b["u_right"][s_, :n].view(np.int32))
                assert np.array_equal(a["depth"][s_, :n].view(np.int32), b["depth"][s_, :n].view(np.int32))

    g = [run(pairs[i], cfg.mb, cfg.mbf) for i in (0, 1, 0)]                 # capture, replay with other images, replay
    g2 = run(list(pairs[1]) + list(pairs[2]), cfg.mb, cfg.mbf)               # 2 pairs, still one chunk: re-capture
    g3 = run(pairs[2], cfg.mb * 0.5, cfg.mbf * 2.0)                          # other matcher parameters: re-capture
    monkeypatch.setenv("JSFE_NO_GRAPH", "1")
    s = [run(pairs[i], cfg.mb, cfg.mbf) for i in (0, 1)]
    s2 = run(list(pairs[1]) + list(pairs[2]), cfg.mb, cfg.mbf)
    s3 = run(pairs[2], cfg.mb * 0.5, cfg.mbf * 2.0)
    same(g[0], s[0]); same(g[1], s[1]); same(g[2], s[0]); same(g2, s2); same(g3, s3)
    assert not np.array_equal(g[0]["kps"], g[1]["kps"])
    # and against the oracle
    ol, orr, (kl, dl, kr, dr), st = _oracle_pair(cfg, *pairs[1])
    n = g[1]["n"][0]
    assert np.array_equal(g[1]["kps"][0, :, :n], kl) and np.array_equal(g[1]["u_right"][0, :n].view(np.int32), st[0].view(np.int32))
    fe.close()


def test_begin_end_two_handles_in_flight_equal_the_blocking_call():
    cfg = CONFIGS["C1"]
    batches = [np.stack([im for s_ in range(3) for im in synth.stereo_pair(cfg.height, cfg.width, 80 + 3 * b + s_)]) for b in range(3)]
    fes = [frontend.Frontend(**cfg.extractor_kwargs(), max_images=6) for _ in range(2)]
    want = []
    for b in batches:
        e = fes[0].process_host_pairs(b, cfg.mb, cfg.mbf)
        want.append({k: np.array(v) for k, v in e.items() if k != "bytes"})
    got = []
    fes[0].process_host_pairs_begin(batches[0], cfg.mb, cfg.mbf, chunk_pairs=3)
    with pytest.raises(frontend.JsfeError):
        fes[0].process_host_pairs_begin(batches[1], cfg.mb, cfg.mbf)          # one batch per handle
    for k in (1, 2):
        fes[k % 2].process_host_pairs_begin(batches[k], cfg.mb, cfg.mbf, chunk_pairs=2 if k == 1 else 3)
        e = fes[(k - 1) % 2].process_host_pairs_end()
        got.append({kk: np.array(v) for kk, v in e.items() if kk != "bytes"})
    e = fes[0].process_host_pairs_end()
    got.append({kk: np.array(v) for kk, v in e.items() if kk != "bytes"})
    with pytest.raises(frontend.JsfeError):
        fes[0].process_host_pairs_end()                                          # nothing in flight
    for g, w in zip(got, want):
        assert np.array_equal(g["n"], w["n"])
        for s_ in range(6):
            n = w["n"][s_]
            assert np.array_equal(g["kps"][s_, :, :n], w["kps"][s_, :, :n]) and np.array_equal(g["desc"][s_, :n], w["desc"][s_, :n])
            if s_ % 2 == 0:
                assert np.array_equal(g["u_right"][s_, :n].view(np.int32), w["u_right"][s_, :n].view(np.int32))
                assert np.array_equal(g["depth"][s_, :n].view(np.int32), w["depth"][s_, :n].view(np.int32))
    for f in fes:
        f.close()


def test_mask_matches_oracle():
    cfg = CONFIGS["tiny"]
    img, _ = synth.stereo_pair(cfg.height, cfg.width, 5)
    mask = np.full((cfg.height, cfg.width), 255, np.uint8)
    mask[:, : cfg.width // 2] = 0
    mask[30:50, :] = 5
    k0, d0 = orc.Oracle(**cfg.extractor_kwargs(), mask=mask).extract(img)
    fe = frontend.Frontend(**cfg.extractor_kwargs(), mask=mask, max_images=1)
    fe.set_images(img)
    fe.extract(0, 1)
    k1, d1 = fe.get_keypoints(0)
    assert np.array_equal(k0, k1) and np.array_equal(d0, d1)


def test_full_size_properties():
    """Size-independent properties at BASELINE's full size (C2), 4 different pairs in one batch."""
    cfg = CONFIGS["C2"]
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=8)
    pairs = [synth.stereo_pair(cfg.height, cfg.width, 10 + s) for s in range(4)]
    fe.set_images(np.stack([im for p in pairs for im in p]))
    fe.extract(0, 8)
    fe.stereo_match(cfg.mb, cfg.mbf, 0, 4)
    for p in range(4):
        k, d = fe.get_keypoints(2 * p)
        n = k.shape[1]
        assert 0 < n <= fe.max_kp
        assert (np.diff(k[4]) >= 0).all(), "keypoints are level-major"
        ang = k[3].view(np.float32)
        assert (ang > -180.0001).all() and (ang <= 180.0).all()
        assert (k[2] > 0).all() and (k[2] <= 4080).all()
        sc = fe.scale[k[4]]
        assert (k[0] >= np.floor(20 * sc) - 1).all() and (k[0] < cfg.width).all()
        ur, dp, bi, bd = fe.get_stereo(p)
        m = ur >= 0
        assert m.sum() > 50
        disp = k[0][m].astype(np.float32) - ur[m]
        assert (disp > 0).all() and (disp < cfg.mbf / cfg.mb).all()
        assert np.allclose(dp[m], cfg.mbf / disp, rtol=1e-6)
        assert (bd[m] < 75).all() and (bi[m] >= 0).all()
        # idempotence: same slot again gives identical bytes
    k_a, d_a = fe.get_keypoints(0)
    fe.extract(0, 8)
    k_b, d_b = fe.get_keypoints(0)
    assert np.array_equal(k_a, k_b) and np.array_equal(d_a, d_b)
