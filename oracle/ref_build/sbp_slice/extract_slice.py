#!/usr/bin/env python
"""Cuts the host code of ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) and of the Frame grid functions it calls
out of the reference checkout BY LINE RANGE, at build time, into oracle/_ref/gen/*.inc (git-ignored).  Nothing of the reference is
committed: this script only names the ranges and checks an anchor string on the first line of each, so that a different
checkout fails loudly instead of compiling something else.

  ORBmatcher.cpp  1647-1963  the live (use_gpu_) branch of SearchByProjection up to the closing brace of the function
  ORBmatcher.cpp  2097-2138  ORBmatcher::ComputeThreeMaxima
  Frame.cpp        464- 479  Frame::AssignFeaturesToGrid
  Frame.cpp        569- 639  Frame::GetFeaturesInArea (the overload SearchByProjection calls)
  Frame.cpp        696- 706  Frame::PosInGrid
"""
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "_ref", "gen")

SLICES = [
    ("src/ORBmatcher.cpp", 1647, 1963, "else", "sbp_live_branch.inc"),
    ("src/ORBmatcher.cpp", 2097, 2138, "void ORBmatcher::ComputeThreeMaxima", "sbp_three_maxima.inc"),
    ("src/Frame.cpp", 464, 479, "void Frame::AssignFeaturesToGrid", "frame_assign_grid.inc"),
    ("src/Frame.cpp", 569, 639, "void Frame::GetFeaturesInArea", "frame_features_in_area.inc"),
    ("src/Frame.cpp", 696, 706, "bool Frame::PosInGrid", "frame_pos_in_grid.inc"),
]


def main():
    os.makedirs(OUT, exist_ok=True)
    for rel, a, b, anchor, name in SLICES:
        lines = open(os.path.join(REF, rel), encoding="utf-8", errors="replace").read().split("\n")
        first = lines[a - 1].strip()
        if not first.startswith(anchor):
            raise SystemExit(f"{rel}:{a} is {first!r}, expected it to start with {anchor!r}: not the reference revision this slice was written for")
        with open(os.path.join(OUT, name), "w") as f:
            f.write(f"// generated from {rel}:{a}-{b} of the reference checkout; do not commit\n")
            f.write("\n".join(lines[a - 1:b]) + "\n")
    print("slices written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
