import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, numpy as np, torch
from jetson_slam_b200 import frontend, synth
from jetson_slam_b200.configs import CONFIGS
cfg = CONFIGS['C2']
fe1 = frontend.Frontend(**cfg.extractor_kwargs(), device=0, max_images=2)
one = torch.empty((2, cfg.height, cfg.width), dtype=torch.uint8).pin_memory()
p = synth.stereo_pair(cfg.height, cfg.width, 0)
one.numpy()[0], one.numpy()[1] = p
for _ in range(50):
    fe1.process_host_pairs(one.numpy(), cfg.mb, cfg.mbf, chunk_pairs=1)
ts = []
for _ in range(500):
    t0 = time.perf_counter(); fe1.process_host_pairs(one.numpy(), cfg.mb, cfg.mbf, chunk_pairs=1); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print("latency ms median %.4f p95 %.4f min %.4f" % (np.median(ts), np.percentile(ts, 95), ts.min()))
