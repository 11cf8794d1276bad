"""Workload configurations (BASELINE.json `configs`).

The reference has no `nFeatures`: the keypoint budget is one candidate per NMS tile per level
(src/cuda/orb_gpu.cpp:305-327), so "N features" is realised by the tile size below.
Common defaults follow the shipped YAMLs (Examples/Stereo/EuRoC.yaml:96-115): scaleFactor 1.2,
th_FAST_MAX 20, FAST_N 9..14, full mask.  `fx`/`bf` are a KITTI-like synthetic calibration
(mbf = bf, mb = bf / fx as Frame.cpp:247 intends).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class FrontendConfig:
    name: str
    height: int
    width: int
    n_levels: int = 8
    scale_factor: float = 1.2
    fast_n_min: int = 9
    fast_n_max: int = 14
    th_fast_min: int = 7
    th_fast_max: int = 20
    tile_h: int = 30
    tile_w: int = 30
    fixed_multi_scale_tile_size: int = 0
    apply_nms_ms: int = 0
    nms_ms_mode_gpu: int = 1
    fx: float = 718.856
    bf: float = 386.1448

    @property
    def mbf(self) -> float:
        return self.bf

    @property
    def mb(self) -> float:
        import numpy as np
        return float(np.float32(self.bf) / np.float32(self.fx))

    def extractor_kwargs(self) -> dict:
        d = asdict(self)
        for k in ("name", "fx", "bf"):
            d.pop(k)
        return d


CONFIGS = {
    # BASELINE.json configs[0..4]
    "C1": FrontendConfig("C1 mono 320x240 ~500 feat", 240, 320, tile_h=38, tile_w=38, fx=277.0, bf=30.0),
    "C2": FrontendConfig("C2 KITTI stereo 1241x376 ~2000 feat L8", 376, 1241, tile_h=46, tile_w=46),
    "C3": FrontendConfig("C3 EuRoC stereo 752x480 ~1000 feat L8", 480, 752, tile_h=56, tile_w=56, fx=435.2, bf=47.9),
    "C4": FrontendConfig("C4 KAIST-VIO 640x480 L4 tile20", 480, 640, n_levels=4, tile_h=20, tile_w=20, fx=380.0, bf=19.0),
    "C5": FrontendConfig("C5 synthetic 1920x1080 ~4000 feat L8", 1080, 1920, tile_h=67, tile_w=67, fx=1100.0, bf=600.0),
    # shipped YAML parameter sets (parity only)
    "KITTI00-02": FrontendConfig("KITTI00-02.yaml L1 tile25 th60", 376, 1241, n_levels=1, th_fast_max=60,
                                 tile_h=25, tile_w=25, apply_nms_ms=1),
    "KITTI04-12": FrontendConfig("KITTI04-12.yaml L8 tile30 th40 nms_ms(gpu)", 376, 1241, th_fast_max=40,
                                 tile_h=30, tile_w=30, apply_nms_ms=1, nms_ms_mode_gpu=1),
    "EuRoC": FrontendConfig("EuRoC.yaml L8 tile30 th20", 480, 752, tile_h=30, tile_w=30, fx=435.2, bf=47.9),
    "KAIST-nmsms-cpu": FrontendConfig("kaist yaml L4 tile20, nms_ms CPU mode", 480, 640, n_levels=4, tile_h=20,
                                      tile_w=20, apply_nms_ms=1, nms_ms_mode_gpu=0, fx=380.0, bf=19.0),
    # small cases for fast CPU tests
    "tiny": FrontendConfig("tiny 160x120 L3 tile16", 120, 160, n_levels=3, tile_h=16, tile_w=16, fx=150.0, bf=12.0),
    "tiny-fixed": FrontendConfig("tiny fixed tiles", 120, 160, n_levels=3, tile_h=13, tile_w=17,
                                 fixed_multi_scale_tile_size=1, fx=150.0, bf=12.0),
}
