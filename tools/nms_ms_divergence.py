#!/usr/bin/env python
"""the reference's GPU cross-scale NMS zeroes s0 entries that other blocks are still reading
(src/cuda/orb_FAST_apply_NMS_MS.cu:235), so its output can depend on scheduling.  This repo implements the deterministic two-phase
rule (all reads see pre-zero values).  Count, over many seeded images, how often the reference's own kernels (oracle/_ref/libjsref.so,
run here on the GPU) differ from that rule (the CPU oracle), and how often two runs of the reference differ from each other.
usage: python tools/nms_ms_divergence.py [n_seeds=24] > profiles/r02_nms_ms_divergence.json"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    from jetson_slam_b200 import frontend, synth
    from jetson_slam_b200.configs import CONFIGS
    from oracle import oracle as orc, ref
    cfg = CONFIGS["KITTI04-12"]       # the shipped YAML that turns the GPU cross-scale rule on
    out = {"config": cfg.name, "seeds": n_seeds, "per_seed": []}
    tot = dict(ref_vs_rule_keypoints=0, ref_run_to_run_keypoints=0, ours_vs_rule_keypoints=0, keypoints=0)
    r = ref.RefEye(**cfg.extractor_kwargs())
    o = orc.Oracle(**cfg.extractor_kwargs())
    fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=1)
    for seed in range(n_seeds):
        img = synth.stereo_pair(cfg.height, cfg.width, 2000 + seed)[0]
        ko, _ = o.extract(img)
        ka, _ = r.extract(img)
        kb, _ = r.extract(img)
        fe.set_images(img[None])
        fe.extract(0, 1)
        kf, _ = fe.get_keypoints(0)

        def diff(a, b):      # keypoints present in one set and not in the other (x, y, octave)
            sa = {tuple(v) for v in a[[0, 1, 4]].T.tolist()}
            sb = {tuple(v) for v in b[[0, 1, 4]].T.tolist()}
            return len(sa ^ sb)
        e = dict(seed=2000 + seed, keypoints=int(ko.shape[1]), ref_vs_rule=diff(ka, ko), ref_run_to_run=diff(ka, kb), ours_vs_rule=diff(kf, ko))
        out["per_seed"].append(e)
        tot["ref_vs_rule_keypoints"] += e["ref_vs_rule"]
        tot["ref_run_to_run_keypoints"] += e["ref_run_to_run"]
        tot["ours_vs_rule_keypoints"] += e["ours_vs_rule"]
        tot["keypoints"] += e["keypoints"]
    out["total"] = tot
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
