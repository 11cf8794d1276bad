#!/usr/bin/env python
"""Print the SASS instructions (with executed counts) that an .ncu-rep attributes to given source lines.
usage: python tools/ncu_sass_of_line.py rep.ncu-rep 331 [332 ...]"""
import csv, io, subprocess, sys

rep, want = sys.argv[1], set(int(a) for a in sys.argv[2:])
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, cur, fname = None, None, None
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if len(r) > 8 and r[0] == 'Line No':
        h = r
        ii = h.index('Instructions Executed')
        continue
    if h is None or len(r) <= ii:
        continue
    if r[0].strip():
        try:
            cur = int(r[0])
        except ValueError:
            cur = None
        if cur in want and fname and fname.endswith('.cuh'):
            print(f'--- L{cur}: {r[1].strip()[:120]}')
        continue
    if cur in want and fname and fname.endswith('.cuh') and r[2].strip():
        print(f'    {r[ii]:>10s}  {r[2].strip()}')
