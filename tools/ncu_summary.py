#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): headline metrics + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--lines 25] [--by-inst]"""
import csv, subprocess, sys, io, collections

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__grid_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']


def main():
    rep = sys.argv[1]
    nlines = int(sys.argv[sys.argv.index('--lines') + 1]) if '--lines' in sys.argv else 25
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit, vals = rows[0], rows[1], rows[-1]
    print('kernel:', vals[hdr.index('Kernel Name')][:80])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'{w:88s} {vals[i]:>18s} {unit[i]}')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next((i for i, r in enumerate(rows) if len(r) > 8 and r[0] == 'Line No'), None)
    if hi is None:
        print('(no source page)'); return
    h = rows[hi]
    ci = h.index('# Samples'); ii = h.index('Instructions Executed'); ti = h.index('Thread Instructions Executed')
    per = collections.OrderedDict()
    cur = None
    for r in rows[hi + 1:]:
        if len(r) < ti + 1:
            continue
        if r[0].strip():
            cur = (r[0], r[1])
            per.setdefault(cur, [0, 0, 0])
        if cur is None or not r[2].strip():
            continue
        try:
            per[cur][0] += int(r[ci] or 0); per[cur][1] += int(r[ii] or 0); per[cur][2] += int(r[ti] or 0)
        except ValueError:
            pass
    tot = sum(v[0] for v in per.values()) or 1
    toti = sum(v[1] for v in per.values()) or 1
    print(f'--- hottest source lines (of {tot} stall samples, {toti} warp-instructions)')
    key = 1 if '--by-inst' in sys.argv else 0
    for (ln, txt), v in sorted(per.items(), key=lambda kv: -kv[1][key])[:nlines]:
        thr = v[2] / v[1] if v[1] else 0
        print(f'{100*v[0]/tot:5.1f}% smp {100*v[1]/toti:5.1f}% inst thr/inst {thr:4.1f}  L{ln:>4s}  {txt.strip()[:105]}')


if __name__ == '__main__':
    main()
