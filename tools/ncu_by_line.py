#!/usr/bin/env python
"""Per-source-line instruction / stall-sample breakdown of one ncu capture.
usage: ncu_by_line.py <report.ncu-rep> <kernel-symbol-substring> [file-substring] [top]"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern_key = sys.argv[1:3]
fkey = sys.argv[3] if len(sys.argv) > 3 else ""
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "bore_b200/lib/libbore_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
sass = start = None
for cubin in sorted(os.listdir(tmp)):  # the cubin that holds the kernel
    txt = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    hit = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern_key in l]
    if hit:
        sass, start = txt, hit[0]
        break
cur, off2line, off2op = None, {}, {}
for l in sass[start + 1:]:
    if l.startswith("//-----"): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: off2line[int(m.group(1), 16)] = cur; off2op[int(m.group(1), 16)] = m.group(2)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[h]
ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
stalls = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
base = int(rows[h + 1][ia], 16)
agg = collections.defaultdict(lambda: [0, 0])
ti = ts = 0
for r in rows[h + 1:]:
    if len(r) <= ie: continue
    k = off2line.get(int(r[ia], 16) - base)
    n = int(r[ie]); s = sum(int(r[i]) for i in stalls)
    ti += n; ts += s
    agg[k][0] += n; agg[k][1] += s
print(f"total inst {ti} samples {ts}")
items = [(k, v) for k, v in agg.items() if k and fkey in k[0]]
for k, v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0]}:{k[1]:5d} inst {v[0]:10d} ({v[0]/ti:6.3f})  samples {v[1]/ts:6.3f}")
