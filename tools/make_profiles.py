#!/usr/bin/env python
"""Turn the .ncu-rep captures of one tag (gpurun_out/prof_<tag>_<kernel>.ncu-rep, taken with `bench.py --pairs 32`)
into the committed evidence: profiles/<round>_ncu_<tag>_<kernel>.txt (tools/ncu_summary.py output) and profiles/traffic.json
(dram__bytes_read.sum + dram__bytes_write.sum per image / per pair, which bench.py scales to its launch size).
usage: python tools/make_profiles.py v7 [pairs_per_launch=32] [round prefix, default r01]"""
import csv, glob, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rnd = sys.argv[3] if len(sys.argv) > 3 else "r01"
out = {}
tj = os.path.join(ROOT, "profiles", "traffic.json")
if os.path.exists(tj):
    out = json.load(open(tj))
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{tag}_*.ncu-rep"))):
    k = os.path.basename(rep)[len(f"prof_{tag}_"):-len(".ncu-rep")]
    txt = os.path.join(ROOT, "profiles", f"{rnd}_ncu_{tag}_{k}.txt")
    s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "--lines", "30"], capture_output=True, text=True).stdout
    open(txt, "w").write(s)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit, vals = rows[0], rows[1], rows[-1]

    def get(name):
        i = hdr.index(name)
        v = float(vals[i].replace(",", ""))
        u = unit[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)

    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    units = pairs if k.startswith("k_stereo") else 2 * pairs
    out[k] = {"dram_bytes_per_unit": (rd + wr) / units, "unit": "pair" if k.startswith("k_stereo") else "image",
              "capture": f"profiles/{rnd}_ncu_{tag}_{k}.txt (ncu --set full, --pairs {pairs}: {units} units per launch)",
              "dram_bytes_read": rd, "dram_bytes_write": wr}
    print(k, "dram MB", (rd + wr) / 1e6, "->", os.path.relpath(txt, ROOT))
json.dump(out, open(tj, "w"), indent=1)
