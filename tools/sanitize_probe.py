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
