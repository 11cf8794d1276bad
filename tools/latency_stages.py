"""Per-kernel CUDA-event times of ONE stereo pair (the latency path) on the C2 workload; run on a GPU box.
usage: python tools/latency_stages.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from jetson_slam_b200 import frontend, synth
from jetson_slam_b200.configs import CONFIGS

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = CONFIGS["C2"]
imgs = np.stack([im for s in range(n_pairs) for im in synth.stereo_pair(cfg.height, cfg.width, s)])
fe = frontend.Frontend(**cfg.extractor_kwargs(), max_images=2 * n_pairs)
fe.set_images(imgs)
for _ in range(5):
    fe.extract(0, 2 * n_pairs); fe.stereo_match(cfg.mb, cfg.mbf, 0, n_pairs)
torch.cuda.synchronize()
fe.profile(True)
fe.profile_read()
N = 100
for _ in range(N):
    fe.extract(0, 2 * n_pairs); fe.stereo_match(cfg.mb, cfg.mbf, 0, n_pairs)
torch.cuda.synchronize()
pr = fe.profile_read()
tot = 0.0
for k, (ms, cnt) in pr.items():
    if cnt:
        print(f"{k:18s} {1e3 * ms / N:8.2f} us per call   ({cnt // N} launches)")
        tot += ms / N
print(f"sum {1e3 * tot:.2f} us (each kernel timed alone between events: no overlap between kernels)")
fe.profile(False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    fe.extract(0, 2 * n_pairs); fe.stereo_match(cfg.mb, cfg.mbf, 0, n_pairs)
e1.record(); torch.cuda.synchronize()
print(f"back to back on the stream (PDL): {1e3 * e0.elapsed_time(e1) / N:.2f} us per pair-set")
