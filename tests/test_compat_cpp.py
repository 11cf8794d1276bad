"""The C++ drop-in boundary: compat/ORBextractor.h + compat/cuda/*.hpp keep the reference's class names and signatures
(include/ORBextractor.h:21-98, include/cuda/orb_gpu.hpp:22-330, include/cuda/synced_mem_holder.hpp:10-65) on top of the
C ABI.  tests/cpp/frame_hotpath.cpp replays Frame::Frame's hot path (two extractor threads, D2H, unpack, stereo match)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from jetson_slam_b200 import synth
from jetson_slam_b200.configs import CONFIGS
from oracle import oracle as orc

DRV = os.path.join(ROOT, "tests", "cpp", "frame_hotpath")


def test_compat_driver_builds_and_links_against_the_c_abi():
    assert os.path.exists(DRV), "tests/cpp/frame_hotpath was not built by __graft_entry__.build()"
    out = subprocess.run(["ldd", DRV], capture_output=True, text=True).stdout
    assert "libjsfe.so" in out and "not found" not in out.split("libjsfe.so")[1].split("\n")[0]


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["C1", "C2"])
def test_frame_hotpath_through_cpp_shims_matches_oracle(name, tmp_path):
    cfg = CONFIGS[name]
    L, R = synth.stereo_pair(cfg.height, cfg.width, 21)
    L.tofile(tmp_path / "l.raw")
    R.tofile(tmp_path / "r.raw")
    out = tmp_path / "out.bin"
    cmd = [DRV, cfg.height, cfg.width, cfg.n_levels, cfg.scale_factor, cfg.fast_n_min, cfg.fast_n_max, cfg.th_fast_min,
           cfg.th_fast_max, cfg.tile_h, cfg.tile_w, tmp_path / "l.raw", tmp_path / "r.raw", repr(cfg.mb), repr(cfg.mbf), out]
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stderr + r.stdout
    raw = np.fromfile(out, np.uint8)
    nl, nr = np.frombuffer(raw[:8], np.int32)
    o = 8
    kl = np.frombuffer(raw[o:o + 24 * nl], np.int32).reshape(6, nl); o += 24 * nl
    dl = raw[o:o + 32 * nl].reshape(nl, 32); o += 32 * nl
    kr = np.frombuffer(raw[o:o + 24 * nr], np.int32).reshape(6, nr); o += 24 * nr
    dr = raw[o:o + 32 * nr].reshape(nr, 32); o += 32 * nr
    ur = np.frombuffer(raw[o:o + 4 * nl], np.float32); o += 4 * nl
    dp = np.frombuffer(raw[o:o + 4 * nl], np.float32)
    ol, orr = orc.Oracle(**cfg.extractor_kwargs()), orc.Oracle(**cfg.extractor_kwargs())
    okl, odl = ol.extract(L)
    okr, odr = orr.extract(R)
    our, odp, _, _ = orc.stereo_match(ol, orr, okl, odl, okr, odr, cfg.mb, cfg.mbf)
    assert np.array_equal(kl, okl) and np.array_equal(kr, okr)
    assert np.array_equal(dl, odl) and np.array_equal(dr, odr)
    assert np.array_equal(ur.view(np.int32), our.view(np.int32)) and np.array_equal(dp.view(np.int32), odp.view(np.int32))


C_BIN = os.path.join(ROOT, "tests", "c", "pair_from_c")


def test_plain_c_consumer_builds():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(C_BIN)


@pytest.mark.gpu
def test_plain_c_consumer_matches_the_python_binding(tmp_path):
    """The same pair through a C99 program that includes only include/jsfe.h and through the ctypes binding."""
    from jetson_slam_b200 import frontend, synth
    from jetson_slam_b200.configs import CONFIGS
    import __graft_entry__ as g
    g.build()
    cfg = CONFIGS["C1"]
    L, R = synth.stereo_pair(cfg.height, cfg.width, 11)
    raw = tmp_path / "pair.raw"
    raw.write_bytes(L.tobytes() + R.tobytes())
    out = subprocess.run([C_BIN, str(cfg.height), str(cfg.width), str(cfg.n_levels), str(cfg.tile_h), str(raw), repr(cfg.mb), repr(cfg.mbf)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    nl, nr, nd, chk = (int(v) for v in out.stdout.split())
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=2)
    r = fe.process_host_pairs(np.stack([L, R]), cfg.mb, cfg.mbf)
    assert (nl, nr) == (int(r["n"][0]), int(r["n"][1])) and nd == int((r["depth"][0, :nl] > 0).sum())
    s = 0
    for i in range(nl):
        for pl in range(6):
            s = (s * 1000003 + (int(r["kps"][0, pl, i]) & 0xFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF
    assert s == chk
    fe.close()
