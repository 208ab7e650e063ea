#!/usr/bin/env python
"""Attribute executed instructions (and warp-stall samples) of an .ncu-rep kernel to source lines (via nvdisasm -g on
the built library).  Usage: python scripts/ncu_lines.py <report.ncu-rep> <mangled kernel name substring> [pixels] [top] [stall]
With a 5th argument the lines are ranked by not-issued stall samples and the dominant stall reasons are shown."""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kname = sys.argv[1], sys.argv[2]
px = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
by_stall = len(sys.argv) > 5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 't2onet_b200/lib/libt2o_b200.so')], cwd=tmp, capture_output=True)
line_of = {}
for f in os.listdir(tmp):
    if not f.endswith('.cubin'):
        continue
    dis = subprocess.run(['nvdisasm', '-g', os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, loc, on = None, None, False
    for ln in dis.splitlines():
        m = re.match(r'\.text\.(\S+):', ln)
        if m:
            on = kname in m.group(1); loc = None
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            loc = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s+/\*([0-9a-f]+)\*/', ln)
        if m:
            line_of[int(m.group(1), 16)] = loc
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix, ia, isrc = h.index('Instructions Executed'), h.index('Address'), h.index('Source')
base = int(rows[2][ia], 16)
by, tot = collections.Counter(), 0
byop = collections.defaultdict(collections.Counter)
ins = h.index('Warp Stall Sampling (Not-issued Samples)')
stall_cols = [(i, c[len('stall_'):-len(' (Not Issued)')]) for i, c in enumerate(h) if c.startswith('stall_') and c.endswith('(Not Issued)')]
st_by, st_tot = collections.Counter(), 0
st_reason = collections.defaultdict(collections.Counter)
for r in rows[2:]:
    if len(r) <= ix:
        continue
    n = int(r[ix]); off = int(r[ia], 16) - base
    loc = line_of.get(off)
    by[loc] += n; tot += n
    st_by[loc] += int(r[ins]); st_tot += int(r[ins])
    for i, name in stall_cols:
        if int(r[i]): st_reason[loc][name] += int(r[i])
    mm = re.match(r'\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)', r[isrc])
    byop[loc][mm.group(2) if mm else '?'] += n
srccache = {}
def text(loc):
    if not loc: return ''
    for d in ('t2onet_b200/csrc', 'include'):
        p = os.path.join(ROOT, d, loc[0])
        if os.path.exists(p):
            if p not in srccache: srccache[p] = open(p).read().splitlines()
            L = srccache[p]
            return L[loc[1]-1].strip()[:90] if loc[1] <= len(L) else ''
    return ''
print('total %.1f thread-instr/px' % (tot * 32 / px))
byfile = collections.Counter()
for loc, n in by.items(): byfile[loc[0] if loc else None] += n
for f, n in byfile.most_common(): print('  %-28s %7.1f /px' % (f, n*32/px))
for loc, n in by.most_common(top):
    ops = ' '.join('%s:%.0f' % (o, c*32/px) for o, c in byop[loc].most_common(4))
    print('%6.1f  %-24s %-90s | %s' % (n*32/px, '%s:%d' % loc if loc else '?', text(loc), ops))

if by_stall:
    print('---- lines by not-issued stall samples (total %d)' % st_tot)
    allr = collections.Counter()
    for loc in st_reason:
        allr.update(st_reason[loc])
    print('   reasons: ' + ' '.join('%s:%.1f%%' % (k, 100 * v / st_tot) for k, v in allr.most_common(10)))
    for loc, n in st_by.most_common(top):
        rs = ' '.join('%s:%.0f%%' % (k, 100 * v / max(n, 1)) for k, v in st_reason[loc].most_common(3))
        print('%5.1f%%  %6.1f i/px  %-24s %-80s | %s' % (100 * n / st_tot, by[loc] * 32 / px, '%s:%d' % loc if loc else '?', text(loc)[:80], rs))
