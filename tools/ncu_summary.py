#!/usr/bin/env python
"""Key metrics, stall mix and executed-instruction mix of every kernel in an .ncu-rep (what profiles/*.txt hold).
usage: ncu_summary.py <report.ncu-rep> [kernel-regex]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "."
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.avg.per_second']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '-k', 'regex:' + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('kernel:', r[hdr.index('Kernel Name')])
    for k in KEYS:
        if k in hdr:
            print(f'  {k:72s} {r[hdr.index(k)]} {units[hdr.index(k)]}')
    st = []
    for k in hdr:
        if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio'):
            st.append((float(r[hdr.index(k)] or 0), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    print('  warp stalls per issued instruction (top 8): ' + ', '.join(f'{k} {v:.2f}' for v, k in sorted(st, reverse=True)[:8]))
    print()
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass', '-k', 'regex:' + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name, h = rows[i][1], rows[i + 1]
        si, ei = h.index('Source'), h.index('Instructions Executed')
        cnt, tot, j = collections.Counter(), 0, i + 2
        while j < len(rows) and rows[j] and rows[j][0] != 'Kernel Name':
            try:
                n = int(rows[j][ei])
            except Exception:
                j += 1
                continue
            t = rows[j][si].split()
            op = t[1] if t and t[0].startswith('@') else (t[0] if t else '?')
            cnt[op.split('.')[0]] += n
            tot += n
            j += 1
        print(f'executed warp-instruction mix: {name[:80]}  total {tot}')
        print('  ' + ', '.join(f'{k} {100 * v / tot:.1f}%' for k, v in cnt.most_common(14)))
        i = j
    else:
        i += 1
