#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key raw metrics + dynamic opcode mix from the source page.
Usage: python scripts/ncu_summary.py <report.ncu-rep> [pixels]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum']
for k in keys:
    if k in m:
        print('%-70s %s' % (k, m[k]))
print('-- stalls per issue')
for h, v in zip(hdr, vals):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and float(v) > 0.05:
        print('   %-40s %s' % (h.split('issue_stalled_')[1].split('_per_issue')[0], v))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix, isrc = h.index('Instructions Executed'), h.index('Source')
by, tot = collections.Counter(), 0
for r in rows[2:]:
    if len(r) <= ix:
        continue
    n = int(r[ix])
    mm = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[isrc])
    by[mm.group(2) if mm else '?'] += n
    tot += n
print('-- warp instructions %d%s' % (tot, '  = %.1f thread-instr / px' % (tot * 32 / px) if px else ''))
for op, n in by.most_common(28):
    print('   %-10s %5.1f%%%s' % (op, 100 * n / tot, '  %7.1f /px' % (n * 32 / px) if px else ''))
