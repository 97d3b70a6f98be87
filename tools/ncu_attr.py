#!/usr/bin/env python
"""Attribute an ncu capture of a kernel to source functions/lines (run on the CPU box).
usage: ncu_attr.py <report.ncu-rep> <cubin-name-substring> <kernel-symbol-substring> [header-for-function-map]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, cubin_key, kern_key = sys.argv[1:4]
hdr_path = sys.argv[4] if len(sys.argv) > 4 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "bore_b200/lib/libbore_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if cubin_key in f][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern_key in l][0]
cur, off2line = None, {}
for l in sass[start + 1:]:
    if l.startswith("//-----"): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: off2line[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[h]; ia, ie, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[h + 1][ia], 16)
byline, sampline = collections.Counter(), collections.Counter()
for r in rows[h + 1:]:
    if len(r) <= isamp: continue
    c = off2line.get(int(r[ia], 16) - base)
    byline[c] += int(r[ie]); sampline[c] += int(r[isamp])
tot, tots = sum(byline.values()), sum(sampline.values())
funcs, src = [], []
if hdr_path:
    src = open(hdr_path).read().split("\n")
    for i, l in enumerate(src, 1):
        m = re.match(r"(?:LB_HD|LB_FN|LB_NI|__device__|template).*?\b(\w+)\(", l)
        if m and not l.startswith(" "): funcs.append((i, m.group(1)))
def fn(c):
    if c is None: return "none"
    f, l = c
    if not hdr_path or f != os.path.basename(hdr_path): return f
    name = "?"
    for i, nm in funcs:
        if i <= l: name = nm
    return name
byf, sf = collections.Counter(), collections.Counter()
for c, n in byline.items(): byf[fn(c)] += n
for c, n in sampline.items(): sf[fn(c)] += n
print("total warp-inst", tot, "samples", tots)
for k, v in sf.most_common(25): print(f"{k:24s} inst {byf[k]/tot:6.3f}  samples {v/tots:6.3f}")
print("top lines by samples")
for c, n in sampline.most_common(30):
    txt = src[c[1] - 1].strip()[:100] if (c and hdr_path and c[0] == os.path.basename(hdr_path)) else ""
    print(c, f"{n/tots:.3f} {byline[c]/tot:.3f}", txt)
