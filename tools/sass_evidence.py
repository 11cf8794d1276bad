#!/usr/bin/env python
"""SASS evidence of the built library (no GPU needed): per kernel the instruction count, the Blackwell/Hopper-specific mnemonics it
contains (TMA: UTMALDG, mbarrier: SYNCS, programmatic dependent launch: griddepcontrol.launch_dependents = PREEXIT, griddepcontrol.wait = ACQBULK, packed-byte ALU:
VABSDIFF4, dot products: IDP, warp reductions: REDUX) and, for k_fast_cells, an opcode histogram per phase (cut at BAR.SYNC).
usage: python tools/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "jetson_slam_b200", "libjsfe.so")
MARK = ["UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "PREEXIT", "ACQBULK", "VABSDIFF4", "IDP", "REDUX", "MATCH", "LDGSTS", "ATOMS", "FLO", "POPC", "PRMT", "LOP3", "IMAD", "HMMA", "UTCHMMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = []
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
        if m and cur:
            kernels[cur].append(m.group(2).strip())
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a)\n")
    print(f"{'kernel':58s} {'SASS':>6s}  mnemonics of interest")
    for (k, ins), name in zip(kernels.items(), demangle):
        ops = collections.Counter()
        for i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") else t[0]
            ops[op.split(".")[0]] += 1
        marks = ", ".join(f"{m} x{ops[m]}" for m in MARK if ops.get(m))
        short = re.sub(r"\(.*", "", name).replace("jsfe::", "").replace("void ", "")
        print(f"{short[:58]:58s} {len(ins):6d}  {marks}")
    # the TMA / mbarrier lines themselves
    print("\n# TMA and mbarrier instructions (what `cp.async.bulk.tensor` / `mbarrier.*` compile to)")
    for (k, ins), name in zip(kernels.items(), demangle):
        hits = [i for i in ins if "UTMALDG" in i or "SYNCS" in i]
        if hits:
            print(re.sub(r"\(.*", "", name).replace("jsfe::", "").replace("void ", ""))
            for h in hits:
                print("    " + h)
    # k_fast_cells<2,false> per phase
    for (k, ins), name in zip(kernels.items(), demangle):
        if "k_fast_cells<2, false>" in name or "k_fast_cells<2, (bool)0>" in name:
            print(f"\n# {name.split('(')[0]}: static opcode histogram per phase (phases are cut at BAR.SYNC: set-up+staging | A compass pre-test + emission | B exact evaluation | C NMS/arg-max | output)")
            phase, hist = 0, collections.Counter()
            for i in ins + ["BAR.SYNC"]:
                t = i.split()
                op = t[1] if t[0].startswith("@") else t[0]
                hist[op.split(".")[0]] += 1
                if "BAR.SYNC" in i:
                    tot = sum(hist.values())
                    print(f"phase {phase}: {tot:5d} instructions: " + ", ".join(f"{o} {c}" for o, c in hist.most_common(10)))
                    phase, hist = phase + 1, collections.Counter()


if __name__ == "__main__":
    main()
