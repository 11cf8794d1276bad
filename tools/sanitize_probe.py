"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): a few configurations incl. cross-scale NMS."""
import sys
import numpy as np
sys.path.insert(0, '.')
from jetson_slam_b200 import frontend, synth
from jetson_slam_b200.configs import CONFIGS
for name in ("tiny", "tiny-fixed", "C1", "KAIST-nmsms-cpu"):
    cfg = CONFIGS[name]
    if name.startswith("KAIST"):
        cfg = cfg.__class__(**{**cfg.__dict__, "height": 200, "width": 260})
    pairs = [synth.stereo_pair(cfg.height, cfg.width, s) for s in range(2)]
    for gpu_mode in ((0, 1) if cfg.apply_nms_ms else (cfg.nms_ms_mode_gpu,)):
        kw = dict(cfg.extractor_kwargs(), nms_ms_mode_gpu=gpu_mode)
        fe = frontend.Frontend(**kw, max_images=4)
        host = np.stack([im for p in pairs for im in p])
        r = fe.process_host_pairs(host, cfg.mb, cfg.mbf, chunk_pairs=1)
        print(name, gpu_mode, r["n"].tolist(), int((r["u_right"] >= 0).sum()))
        fe.close()

# single-pair call (CUDA-graph replay path) + the adjacent rows: frame view, device grid, fused projection search, rectification
import torch
cfg = CONFIGS["C1"]
fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=2)
L, R = synth.stereo_pair(cfg.height, cfg.width, 3)
for _ in range(2):
    r = fe.process_host_pairs(np.stack([L, R]), cfg.mb, cfg.mbf)
print("graph path", r["n"].tolist(), int((r["u_right"] >= 0).sum()))
view = frontend.frame_view(fe, 0)
n = int(r["n"][0])
last, cur, Rm, t = synth.projection_scene(n_cur=700, n_last=600, seed=2)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
dl, dc = {k: d(v) for k, v in last.items()}, {k: d(v) for k, v in cur.items()}
sf = np.cumprod(np.array([1.0] + [1.2] * 7, np.float32)).astype(np.float32)
out = frontend.search_by_projection(dl, dc, d(Rm), d(t), **synth.SBP_K, **synth.SBP_BOUNDS, mbf=synth.SBP_MBF, th=7.0, scale_factors=sf, level_mode=0)
torch.cuda.synchronize()
print("sbp matches", int(out["n_matches"].cpu()[0]), "view x[:3]", view["x"][:3].cpu().tolist(), n)
src = d(np.random.default_rng(0).integers(0, 256, size=(3, 97, 131), dtype=np.uint8))
mx = d(np.random.default_rng(1).uniform(-10, 140, size=(80, 123)).astype(np.float32))
my = d(np.random.default_rng(2).uniform(-10, 105, size=(80, 123)).astype(np.float32))
o = frontend.remap_bilinear(src, mx, my)
g = frontend.cvt_gray(d(np.random.default_rng(3).integers(0, 256, size=(33, 45, 3), dtype=np.uint8)))
torch.cuda.synchronize()
print("remap sum", int(o.sum().cpu()), "gray sum", int(g.sum().cpu()))
fe.close()
