#!/usr/bin/env python
"""Aggregate an .ncu-rep's per-instruction samples over source-line ranges.
usage: python tools/ncu_phases.py rep name:lo-hi [name:lo-hi ...]   (lines of the first CUDA source file in the report)"""
import csv, io, subprocess, collections, sys


def load(rep):
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    per = collections.OrderedDict()
    cur, fname, h = None, None, None
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if len(r) > 8 and r[0] == 'Line No':
            h = r
            ci, ii, ti = h.index('# Samples'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')
            continue
        if h is None or len(r) < ti + 1:
            continue
        if r[0].strip():
            try:
                cur = (fname, int(r[0]))
            except ValueError:
                cur = None
            if cur:
                per.setdefault(cur, [0, 0, 0, r[1]])
            continue
        if cur is None or not r[2].strip():
            continue
        try:
            per[cur][0] += int(r[ci] or 0); per[cur][1] += int(r[ii] or 0); per[cur][2] += int(r[ti] or 0)
        except ValueError:
            pass
    return per


def main():
    per = load(sys.argv[1])
    tot = [sum(v[k] for v in per.values()) for k in range(3)]
    print('total samples', tot[0], 'warp inst', tot[1])
    files = collections.Counter()
    for (f, ln), v in per.items():
        files[f] += v[1]
    print('files:', dict(files))
    main_file = files.most_common(1)[0][0]
    for spec in sys.argv[2:]:
        nm, rng = spec.split(':')
        a, b = map(int, rng.split('-'))
        s = [sum(v[k] for (f, ln), v in per.items() if f == main_file and a <= ln <= b) for k in range(3)]
        if s[1]:
            print(f'{nm:24s} L{a}-{b}: smp {100*s[0]/tot[0]:5.1f}%  inst {100*s[1]/tot[1]:5.1f}%  thr/inst {s[2]/s[1]:5.1f}')
    s = [sum(v[k] for (f, ln), v in per.items() if f != main_file) for k in range(3)]
    if s[1]:
        print(f'{"(other files)":24s}: smp {100*s[0]/tot[0]:5.1f}%  inst {100*s[1]/tot[1]:5.1f}%  thr/inst {s[2]/s[1]:5.1f}')


if __name__ == '__main__':
    main()
