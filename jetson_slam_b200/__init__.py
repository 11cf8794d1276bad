"""jetson_slam_b200: B200-native stereo visual-SLAM front-end (pyramid, FAST/NMS, IC-angle, rBRIEF, stereo match).

The compute path is hand-written sm_100a CUDA behind the C ABI in include/jsfe.h (libjsfe.so, built
in-tree by __graft_entry__.build()).  This package holds the CUDA sources (csrc/), the ctypes host
binding mirroring the reference's ORBExtractor / ORB_GPU interface (frontend.py) and the synthetic
input generator (synth.py).  There is no CPU fallback: importing the binding without the built
library raises.
"""
