#!/usr/bin/env python
"""Per-phase instruction / stall-sample shares of one ncu capture of fit_unit_kernel (fit_unit.cu), from the
per-line output of tools/ncu_by_line.py.  usage: ncu_fit_unit_phases.py <by_line.txt> [steps] [ctas]"""
import os, re, sys
rows = []
for l in open(sys.argv[1]):
    m = re.match(r'\s*(\S+):\s*(\d+) inst\s+(\d+) \(\s*([\d.]+)\)\s+samples\s+([\d.]+)', l)
    if m: rows.append((m.group(1), int(m.group(2)), int(m.group(3)), float(m.group(5))))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 992
ctas = int(sys.argv[3]) if len(sys.argv) > 3 else 8
tot = sum(r[2] for r in rows)
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "bore_b200/csrc/fit_unit.cu")).read().split('\n')
def find(s): return next(i + 1 for i, l in enumerate(src) if s in l)
marks = [('primitives (FFMA2, mbarrier, bulk copy)', find('cluster / packed-FMA primitives')),
         ('fu_partial (GEMM pass, 1st half)', find('void fu_partial(')), ('fu_reduce (GEMM pass, 2nd half)', find('void fu_reduce(')),
         ('fu_grad_tile', find('void fu_grad_tile(')), ('row dot / sum', find('float fu_row_dot(')),
         ('kernel head / staging', find('fit_unit_kernel(const FuArgs a)')), ('gather', find('auto row_of')),
         ('step head', find('int par = 0;           // parity')), ('forward', find('// ---- forward through the hidden layers')),
         ('logit / loss', find('// ---- Dense(1) logit')), ('Adam scalars', find('// ---- Adam scalars of this step')),
         ('delta of the last hidden layer', find('// ---- delta of the last hidden layer')),
         ('reverse', find('// ---- reverse through the hidden layers')), ('gradient items: decode', find('// ---- weight gradients of the slices')),
         ('gradient items: compute', find('        if (kind == 0) {')), ('Adam pass', find('// ---- Adam, in place')),
         ('write-back', find('// ---- write back: the column copies')), ('end', len(src) + 1)]
print(f'warp instructions {tot} = {tot / steps / ctas / 1e3:.1f} k per CTA and step')
for (n, a), (_, b) in zip(marks, marks[1:]):
    i = sum(r[2] for r in rows if r[0] == 'fit_unit.cu' and a <= r[1] < b)
    s = sum(r[3] for r in rows if r[0] == 'fit_unit.cu' and a <= r[1] < b)
    print(f'{n:42s} lines {a:4d}-{b:4d}  instructions {i / tot:.3f} ({i / steps / ctas / 1e3:4.1f} k per CTA-step)  stall samples {s:.3f}')
oth = [r for r in rows if r[0] != 'fit_unit.cu']
print(f'other files (fit_common.cuh, intrinsics): instructions {sum(r[2] for r in oth) / tot:.3f}  stall samples {sum(r[3] for r in oth):.3f}')
print('top lines by stall samples:')
for r in sorted(rows, key=lambda r: -r[3])[:14]:
    print(f'  {r[0]}:{r[1]}  {r[2] / steps / ctas / 1e3:.2f} k inst  samples {r[3]:.3f}  ', src[r[1] - 1].strip()[:100] if r[0] == 'fit_unit.cu' else '')
