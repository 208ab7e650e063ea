#!/usr/bin/env python
"""Split an .ncu-rep kernel's executed instructions and warp-stall samples at its BAR.SYNC instructions (the phases of
the row-pipeline kernels) and list the most-stalled instructions.  Usage: python scripts/ncu_phases.py <report.ncu-rep> [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ia, isrc, ix = h.index('Address'), h.index('Source'), h.index('Instructions Executed')
ins, iall = h.index('Warp Stall Sampling (Not-issued Samples)'), h.index('Warp Stall Sampling (All Samples)')
base = int(rows[2][ia], 16)
rows = [r for r in rows[2:] if len(r) > ix]
segs, cur = [], [0, 0, 0]
for r in rows:
    cur[0] += int(r[ins]); cur[1] += int(r[ix]); cur[2] += int(r[iall])
    if 'BAR.SYNC' in r[isrc]:
        segs.append((int(r[ia], 16) - base, cur)); cur = [0, 0, 0]
segs.append((-1, cur))
tn, ti, ta = (sum(c[i] for _, c in segs) for i in range(3))
print('segment end   not-issued   instr%%   samples%%   (totals: %d not-issued, %d warp-instr, %d samples)' % (tn, ti, ta))
for a, c in segs:
    if c[2] * 200 > ta:
        print('  %8s   %5.1f%%      %5.1f%%    %5.1f%%    not-issued/samples %.2f' % (hex(a) if a >= 0 else 'end', 100 * c[0] / tn, 100 * c[1] / ti, 100 * c[2] / ta, c[0] / max(c[2], 1)))
for t in sorted(((int(r[ins]), int(r[ia], 16) - base, r[isrc]) for r in rows), reverse=True)[:top]:
    print('%6d  %7s  %s' % (t[0], hex(t[1]), t[2].strip()[:100]))
