#!/usr/bin/env python
"""Pin the CPU oracle against the reference itself.  RUNS ON THE GPU BOX (needs oracle/_ref/libjsref.so).

For every (config, seed) it feeds the same seeded synthetic images to
  O1 = the reference's src/cuda compiled unmodified for sm_100a   (oracle/ref.py)
  O2 = the CPU restatement                                         (oracle/oracle.py)
compares every stage bit-for-bit, prints a report, and writes golden vectors of O1's outputs to
gpurun_out/golden/*.npz (copy them into tests/golden/ and commit).  Large intermediates are stored as
sha256 digests; keypoints / descriptors / stereo outputs are stored in full.

Usage: python tools/make_golden.py [--out gpurun_out/golden]
"""
import argparse, hashlib, json, os, sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jetson_slam_b200 import synth
from jetson_slam_b200.configs import CONFIGS
from oracle import oracle as orc
from oracle import ref


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


CASES = [("C1", 0), ("C1", 1), ("C2", 0), ("C2", 1), ("C3", 0), ("C4", 0), ("KITTI00-02", 0), ("EuRoC", 0),
         ("KAIST-nmsms-cpu", 0), ("KITTI04-12", 0), ("tiny", 0), ("tiny", 1), ("tiny-fixed", 0), ("C5", 0)]


def run_case(name, seed, out_dir, report):
    cfg = CONFIGS[name]
    kw = cfg.extractor_kwargs()
    L, R = synth.stereo_pair(cfg.height, cfg.width, seed)
    rl, rr = ref.RefEye(**kw), ref.RefEye(**kw)
    ol, orr = orc.Oracle(**kw), orc.Oracle(**kw)
    res = {"case": f"{name}/seed{seed}"}
    gold = {"img_l_sha": sha(L), "img_r_sha": sha(R)}
    for eye, img, r, o in (("l", L, rl, ol), ("r", R, rr, orr)):
        k1, d1 = r.extract(img)
        k2, d2 = o.extract(img)
        gold[f"kps_{eye}"] = k1
        gold[f"desc_{eye}"] = d1
        x, y, s, a, n, off = r.level_keypoints()
        ox, oy, os_, oa = o.level_keypoints()
        on = o.n_keypoints()
        gold[f"n_per_level_{eye}"] = n
        lv = {"img": 0, "blur": 0, "score": 0}
        img_sha, blur_sha, score_sha = [], [], []
        for l in range(cfg.n_levels):
            a1, a2 = r.level_image(l), o.level_image(l)
            lv["img"] += int((a1 != a2).sum()); img_sha.append(sha(a1))
            b1, b2 = r.level_blur(l), o.level_blur(l)
            lv["blur"] += int((b1 != b2).sum()); blur_sha.append(sha(b1))
            s1, s2 = r.level_score(l), o.level_score(l)
            lv["score"] += int((s1 != s2).sum()); score_sha.append(sha(s1))
        gold[f"img_sha_{eye}"] = np.array(img_sha)
        gold[f"blur_sha_{eye}"] = np.array(blur_sha)
        gold[f"score_sha_{eye}"] = np.array(score_sha)
        same_n = bool((n == on).all())
        lk = {"n_equal": same_n}
        if same_n:
            idx = np.concatenate([off[l] + np.arange(n[l]) for l in range(cfg.n_levels)]).astype(np.int64)
            lk["xy_mismatch"] = int(((x[idx] != ox[idx]) | (y[idx] != oy[idx]) | (s[idx] != os_[idx])).sum())
            lk["angle_bit_mismatch"] = int((a[idx].view(np.int32) != oa[idx].view(np.int32)).sum())
            gold[f"angle_{eye}"] = a[idx]
        res[eye] = {"N_ref": int(k1.shape[1]), "N_orc": int(k2.shape[1]), "level_px_mismatch": lv, "level_kp": lk,
                    "kps_equal": bool(k1.shape == k2.shape and (k1 == k2).all()),
                    "desc_bit_mismatch": int(np.unpackbits(d1 ^ d2).sum()) if d1.shape == d2.shape else -1}
        if eye == "l":
            kl1, dl1, kl2, dl2 = k1, d1, k2, d2
        else:
            kr1, dr1, kr2, dr2 = k1, d1, k2, d2
    if kl1.shape[1] > 0 and kr1.shape[1] > 0:
        ur1, dp1 = ref.stereo_match(rl, rr, kl1.shape[1], cfg.mb, cfg.mbf)
        ur2, dp2, bi2, bd2 = orc.stereo_match(ol, orr, kl1, dl1, kr1, dr1, cfg.mb, cfg.mbf)  # same inputs as O1
        gold["u_right"] = ur1
        gold["depth"] = dp1
        gold["orc_best_idx_r"] = bi2
        gold["orc_best_dist"] = bd2
        res["stereo"] = {"matched_ref": int((ur1 >= 0).sum()), "matched_orc": int((ur2 >= 0).sum()),
                         "mask_equal": bool(((ur1 >= 0) == (ur2 >= 0)).all()),
                         "u_right_maxabs": float(np.abs(ur1 - ur2).max()), "depth_maxabs": float(np.abs(dp1 - dp2).max()),
                         "u_right_bits_equal": bool((ur1.view(np.int32) == ur2.view(np.int32)).all()),
                         "depth_bits_equal": bool((dp1.view(np.int32) == dp2.view(np.int32)).all())}
    gold["mb_mbf"] = np.array([cfg.mb, cfg.mbf], np.float32)
    np.savez_compressed(os.path.join(out_dir, f"ref_{name}_seed{seed}.npz"), **gold)
    rl.close(); rr.close()
    report.append(res)
    print(json.dumps(res), flush=True)


def tables_check(report):
    cfg = CONFIGS["C2"]
    r = ref.RefEye(**cfg.extractor_kwargs())
    o = orc.Oracle(**cfg.extractor_kwargs())
    lut, umax, g, px, py = r.tables()
    opx, opy = o.pattern()
    res = {"case": "tables", "lut_equal": bool((lut != 0).astype(np.uint8).tolist() == o.lut()[:0xFFFF].tolist()),
           "umax_equal": bool((umax == o.umax()).all()),
           "gauss_bits_equal": bool((g.view(np.int32) == o.gauss().view(np.int32)).all()),
           "pattern_equal": bool((px == opx).all() and (py == opy).all())}
    r.close()
    report.append(res)
    print(json.dumps(res), flush=True)


def degenerate_check(out_dir, report):
    cfg = CONFIGS["C1"]
    kw = cfg.extractor_kwargs()
    r = ref.RefEye(**kw)
    o = orc.Oracle(**kw)
    gold = {}
    for nm, img in synth.degenerate_images(cfg.height, cfg.width).items():
        k1, d1 = r.extract(img)
        k2, d2 = o.extract(img)
        gold[f"kps_{nm}"] = k1
        gold[f"desc_{nm}"] = d1
        res = {"case": f"degenerate/{nm}", "N_ref": int(k1.shape[1]), "N_orc": int(k2.shape[1]),
               "kps_equal": bool(k1.shape == k2.shape and (k1 == k2).all()),
               "desc_bit_mismatch": int(np.unpackbits(d1 ^ d2).sum()) if d1.shape == d2.shape else -1}
        report.append(res)
        print(json.dumps(res), flush=True)
    np.savez_compressed(os.path.join(out_dir, "ref_degenerate_C1.npz"), **gold)
    r.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/golden")
    ap.add_argument("--cases", default="")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    report = []
    tables_check(report)
    degenerate_check(args.out, report)
    want = set(args.cases.split(",")) if args.cases else None
    for name, seed in CASES:
        if want and name not in want:
            continue
        run_case(name, seed, args.out, report)
    json.dump(report, open(os.path.join(args.out, "pin_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
