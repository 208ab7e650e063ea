#!/usr/bin/env python
"""Static instruction count of a kernel's SASS between its BAR.SYNC instructions (the phases of the row-pipeline kernels),
by opcode class.  Usage: python scripts/dev/sass_phases.py ['<demangled kernel substring>'] [lib.so]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
want = sys.argv[1] if len(sys.argv) > 1 else 'step_sharp_kernel<4, false, 256, false, 7750433u, 2>'
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, 't2onet_b200', 'lib', 'libt2o_b200.so')
syms = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
mangled = re.findall(r'Function : (\S+)', syms)
dem = subprocess.run(['c++filt'], input='\n'.join(mangled), capture_output=True, text=True).stdout.splitlines()
name = next(m for m, d in zip(mangled, dem) if want in d)
sass = subprocess.run(['cuobjdump', '-sass', '-fun', name, lib], capture_output=True, text=True).stdout
ins = []
for ln in sass.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(@!?U?P\w+\s+)?([A-Z0-9_.]+)(.*?);', ln)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
segs, cur = [], []
for a, op, rest in ins:
    cur.append((a, op, rest))
    if op.startswith('BAR'):
        segs.append(cur); cur = []
segs.append(cur)
FMA = ('FADD', 'FMUL', 'FFMA', 'FFMA2', 'FADD2', 'FMUL2')
ALU = ('FSEL', 'FSETP', 'ISETP', 'IADD3', 'IADD', 'LEA', 'MOV', 'FMNMX', 'FMNMX3', 'LOP3', 'SEL', 'SHF', 'PLOP3', 'IMNMX', 'VIMNMX', 'VIADD', 'IABS', 'FCHK', 'CS2R', 'PRMT', 'SGXT', 'R2P', 'P2R')
print('%d instructions, %d segments' % (len(ins), len(segs)))
for i, sg in enumerate(segs):
    if len(sg) < 40:
        continue
    c = collections.Counter(op.split('.')[0] for _, op, _ in sg)
    fma = sum(v for k, v in c.items() if k in FMA)
    alu = sum(v for k, v in c.items() if k in ALU)
    print('segment %d: 0x%x..0x%x  %d instr  (fma-pipe %d, alu-pipe %d, other %d)' % (i, sg[0][0], sg[-1][0], len(sg), fma, alu, len(sg) - fma - alu))
    print('    ' + '  '.join('%s %d' % kv for kv in c.most_common(24)))
