#!/usr/bin/env python
"""Headline metrics of EVERY kernel in an .ncu-rep that holds several kernels (one `ncu --set full -k regex:a|b|c` run): for each
distinct kernel name the first captured launch.  usage: python tools/ncu_multi_summary.py rep > profiles/xyz.txt"""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_static',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']


def main():
    raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    seen = set()
    for r in rows[2:]:
        name = r[ki]
        short = name.split('(')[0]
        if short in seen:
            continue
        seen.add(short)
        print('kernel:', name[:110])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f'  {w:84s} {r[i]:>16s} {unit[i]}')
        print()


if __name__ == '__main__':
    main()
