#!/usr/bin/env python
"""Regression pins for the adjacent rows: outputs of the oracle restatements on fixed seeded scenes, so that an
accidental change of the oracle (the checker of the CUDA path) is caught on the CPU.  These are NOT reference outputs: the
helper kernels are pinned against the reference's kernels on the GPU box (tests/test_helpers.py), f1's host loops against the
reference's own host code (tools/make_golden_sbp.py -> tests/golden/sbpref_*.npz).
usage: python tools/make_golden_adjacent.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from jetson_slam_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

SF = np.cumprod(np.array([1.0] + [1.2] * 7, np.float32)).astype(np.float32)


def main():
    out = {}
    for mode in (0, 1, 2):
        last, cur, R, t = synth.projection_scene(n_cur=900, n_last=700, seed=40 + mode)
        r = orc.search_by_projection(last, cur, R, t, **synth.SBP_K, **synth.SBP_BOUNDS, mbf=synth.SBP_MBF, th=7.0 if mode != 1 else 15.0,
                                     scale_factors=SF, level_mode=mode)
        for k in ("best_idx2", "best_dist", "rot_bin", "cur_match", "hist"):
            out[f"m{mode}_{k}"] = np.asarray(r[k])
        out[f"m{mode}_nmatches"] = np.array(r["nmatches"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "adj_sbp_oracle.npz"), **out)
    print("written", {k: int(v) for k, v in out.items() if k.endswith("nmatches")})


if __name__ == "__main__":
    main()
