#!/usr/bin/env python
"""Per-operator static instruction cost (scripts/micro/op_cost.cu): the op applied twice in a loop of 2 iterations; prints
(total - identity) / 4 pixels by opcode class."""
import collections, os, re, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
cub = '/tmp/op_cost.cubin'
subprocess.run(['nvcc', '-arch=sm_100a', '-O3', '-cubin', '-o', cub, os.path.join(HERE, 'op_cost.cu')], check=True)
sass = subprocess.run(['cuobjdump', '-sass', cub], capture_output=True, text=True).stdout
funcs = {}
cur = None
for ln in sass.splitlines():
    m = re.search(r'Function : (\S+)', ln)
    if m:
        cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip(); funcs[cur] = collections.Counter(); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)', ln)
    if m and cur:
        funcs[cur][m.group(3).split('.')[0]] += 1
NAMES = {'0': 'brightness', '1': 'contrast', '2': 'saturation', '3': 'color', '5': 'tone', '-1': 'identity'}
base = {}
for k, c in funcs.items():
    m = re.search(r'(bwd_k|fwd_k)<(-?\d+), (true|false)>', k)
    if m and m.group(2) == '-1':
        base[m.group(1)] = c
for k, c in sorted(funcs.items()):
    m = re.search(r'(bwd_k|fwd_k)<(-?\d+), (true|false)>', k)
    if not m or m.group(2) == '-1':
        continue
    d = collections.Counter(c); d.subtract(base[m.group(1)])
    tot = sum(d.values())
    print('%s %-10s CL=%-5s %6.1f instr/px   %s' % (m.group(1), NAMES[m.group(2)], m.group(3), tot / 4.0,
          '  '.join('%s %.1f' % (o, n / 4.0) for o, n in d.most_common(14) if n)))
