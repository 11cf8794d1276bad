#!/usr/bin/env python
"""Launch every kernel that bench.py's main loop does not, a few times each, so that ONE `ncu` run over this script can capture them
(tools/gpu_evidence_r02.sh):
  ours       k_nms_ms_dense / k_nms_ms_buckets (cross-scale NMS configs), k_repitch + the graph path (process_host_pairs),
             k_frame_grid / k_sbp_match / k_sbp_finish, k_project_points, k_hamming_pairs, k_in_frustum, k_frame_view, k_pack,
             k_gather_pack, k_cvt_gray, k_remap_bilinear
  reference  (--ref) the reference's own src/cuda (oracle/_ref/libjsref.so) on the same C2 pair: its FAST score, NMS, pyramid, blur,
             orientation, descriptor and stereo kernels (captures of the corresponding reference kernels)
usage: python tools/exercise_kernels.py [--ref] [--reps 3]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from jetson_slam_b200 import synth
    from jetson_slam_b200.configs import CONFIGS
    cfg = CONFIGS["C2"]
    L, R = synth.stereo_pair(cfg.height, cfg.width, 0)
    if a.ref:
        from oracle import ref
        rl, rr = ref.RefEye(**cfg.extractor_kwargs()), ref.RefEye(**cfg.extractor_kwargs())
        for _ in range(a.reps):
            kl, dl = rl.extract(L)
            kr, dr = rr.extract(R)
            ref.stereo_match(rl, rr, kl.shape[1], cfg.mb, cfg.mbf)
        print("reference kernels launched:", kl.shape[1], kr.shape[1])
        return
    import torch
    from jetson_slam_b200 import distributed as jd, frontend
    for name in ("KITTI04-12", "KAIST-nmsms-cpu"):           # cross-scale NMS, both rules
        c = CONFIGS[name]
        im = synth.stereo_pair(c.height, c.width, 1)
        fe = frontend.Frontend(**c.extractor_kwargs(), max_images=2)
        fe.set_images(np.stack(im))
        for _ in range(a.reps):
            fe.extract(0, 2)
        torch.cuda.synchronize()
        fe.close()
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=8)
    host = np.stack([L, R] * 4)
    for _ in range(a.reps):
        fe.process_host_pairs(host, cfg.mb, cfg.mbf, chunk_pairs=2)      # k_repitch + chain
        fe.process_host_pairs(host[:2], cfg.mb, cfg.mbf, chunk_pairs=1)  # the single-pair graph
        fe.get_keypoints(0)
        view = frontend.frame_view(fe, 0)                                 # k_frame_view
    g = jd.Gatherer(fe, 4)
    for _ in range(a.reps):
        g.begin(0, 4, None)
        g.end()
    g.close()
    # adjacent rows
    last, cur, Rm, t = synth.projection_scene(n_cur=3412, n_last=3412, seed=7)
    d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    dl, dc, dR, dt = {k: d(v) for k, v in last.items()}, {k: d(v) for k, v in cur.items()}, d(Rm), d(t)
    sf = np.cumprod(np.array([1.0] + [1.2] * 7, np.float32)).astype(np.float32)
    kw = dict(**synth.SBP_K, **synth.SBP_BOUNDS, mbf=synth.SBP_MBF, th=7.0, scale_factors=sf, level_mode=0)
    for _ in range(a.reps):
        frontend.search_by_projection(dl, dc, dR, dt, **kw)
    n = 20000
    rng = np.random.default_rng(0)
    P = d(rng.normal(0, 5, size=(3, n)).astype(np.float32) + np.array([[0], [0], [20]], np.float32))
    Pn = d(rng.normal(0, 1, size=(3, n)).astype(np.float32))
    far, near = d(np.full(n, 80.0, np.float32)), d(np.full(n, 0.5, np.float32))
    Ow = d(np.zeros(3, np.float32))
    il = d(rng.integers(0, 3412, size=n).astype(np.int32))
    ir = d(rng.integers(0, 3412, size=n).astype(np.int32))
    K, Bd = synth.SBP_K, synth.SBP_BOUNDS
    for _ in range(a.reps):
        frontend.project_points(P, dR, dt, **K, **Bd)
        frontend.hamming_pairs(il, ir, dl["desc"], dc["desc"])
        frontend.in_frustum(P, Pn, far, far, near, dR, dt, Ow, K["fx"], K["fy"], K["cx"], K["cy"], int(Bd["min_x"]), int(Bd["max_x"]), int(Bd["min_y"]),
                            int(Bd["max_y"]), 8, float(np.log(1.2)), 0.5)
    img = d(rng.integers(0, 256, size=(480, 752, 3), dtype=np.uint8))
    for _ in range(a.reps):
        frontend.cvt_gray(img)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
