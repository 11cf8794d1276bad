#!/usr/bin/env python
"""Per-phase instruction counts of a kernel from an .ncu-rep: the SASS of the profiled kernel is cut at its BAR.SYNC instructions
(phases in program order) and, inside a phase, at user-given source-line ranges; prints warp instructions, thread instructions and
stall samples per piece.  usage: python tools/ncu_sass_phases.py rep [name=lo-hi ...]   (lo-hi: CUDA source lines, inclusive)"""
import csv, io, subprocess, sys, collections


def load(rep):
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None
    cur = None
    sass = {}
    for r in rows:
        if len(r) > 8 and r[0] == 'Line No':
            h = r
            ci, ii, ti = h.index('# Samples'), h.index('Instructions Executed'), h.index('Thread Instructions Executed')
            continue
        if h is None or len(r) < ti + 1:
            continue
        if r[0].strip():
            try:
                cur = int(r[0])
            except ValueError:
                cur = None
            continue
        if cur is None or not r[2].strip() or r[2] == '...':
            continue
        try:
            a = int(r[2], 16)
        except ValueError:
            continue
        if a not in sass:   # an instruction attributed to several inlined lines is listed once per line: keep the innermost (first)
            sass[a] = (cur, r[3].strip(), int(r[ii] or 0), int(r[ci] or 0), int(r[ti] or 0))
    return [(a,) + sass[a] for a in sorted(sass)]


def main():
    rep = sys.argv[1]
    ranges = []
    for spec in sys.argv[2:]:
        nm, rng = spec.split('=')
        lo, hi = map(int, rng.split('-'))
        ranges.append((nm, lo, hi))
    s = load(rep)
    tot_w = sum(x[3] for x in s) or 1
    tot_s = sum(x[4] for x in s) or 1
    tot_t = sum(x[5] for x in s)
    print(f'total: {tot_w} warp-inst, {tot_t} thread-inst, {tot_s} samples, {len(s)} static SASS')
    bars = [i for i, x in enumerate(s) if 'BAR.SYNC' in x[2]]
    cuts = [0] + [b + 1 for b in bars] + [len(s)]
    for k, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        seg = s[a:b]
        w, sm, t = sum(x[3] for x in seg), sum(x[4] for x in seg), sum(x[5] for x in seg)
        print(f'phase {k}: static {len(seg):5d}  warp-inst {100*w/tot_w:5.1f}%  thread-inst {100*t/max(tot_t,1):5.1f}%  samples {100*sm/tot_s:5.1f}%  thr/inst {t/max(w,1):4.1f}')
        for nm, lo, hi in ranges:
            sub = [x for x in seg if lo <= x[1] <= hi]
            if sub:
                w2, sm2, t2 = sum(x[3] for x in sub), sum(x[4] for x in sub), sum(x[5] for x in sub)
                print(f'    {nm:18s} warp-inst {100*w2/tot_w:5.1f}%  thread-inst {100*t2/max(tot_t,1):5.1f}%  samples {100*sm2/tot_s:5.1f}%  thr/inst {t2/max(w2,1):4.1f}')
    mn = collections.Counter()
    for x in s:
        op = x[2].split()
        op = op[1] if op[0].startswith('@') else op[0]
        mn[op.split('.')[0]] += x[3]
    print('top opcodes:', ', '.join(f'{k} {100*v/tot_w:.1f}%' for k, v in mn.most_common(14)))


if __name__ == '__main__':
    main()
