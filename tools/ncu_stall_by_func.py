#!/usr/bin/env python
"""Per-function stall-reason breakdown of one ncu capture (source page + nvdisasm line info).
usage: ncu_stall_by_func.py <report.ncu-rep> <kernel-symbol-substring> <header-for-function-map>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, kern_key, hdr_path = sys.argv[1:4]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "bore_b200/lib/libbore_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if "lbfgsb" in f][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern_key in l][0]
cur, off2line = None, {}
for l in sass[start + 1:]:
    if l.startswith("//-----"): break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: off2line[int(m.group(1), 16)] = cur
src = open(hdr_path).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:LB_HD|LB_FN|LB_NI|__device__|template).*?\b(\w+)\(", l)
    if m and not l.startswith(" "): funcs.append((i, m.group(1)))
def fn(c):
    if c is None: return "none"
    f, l = c
    if f != os.path.basename(hdr_path): return f
    name = "?"
    for i, nm in funcs:
        if i <= l: name = nm
    return name
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[h]
ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
stalls = [(i, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
base = int(rows[h + 1][ia], 16)
agg = collections.defaultdict(lambda: collections.Counter())
for r in rows[h + 1:]:
    if len(r) <= ie: continue
    f = fn(off2line.get(int(r[ia], 16) - base))
    agg[f]["inst"] += int(r[ie])
    for i, c in stalls: agg[f][c] += int(r[i])
tot = sum(sum(v[c] for _, c in stalls) for v in agg.values())
cols = ["stall_no_inst", "stall_wait", "stall_short_sb", "stall_long_sb", "stall_branch_resolving", "stall_selected"]
print(f"{'function':24s} {'inst':>10s} {'samples':>8s}  " + " ".join(f"{c[6:14]:>8s}" for c in cols))
for f, v in sorted(agg.items(), key=lambda kv: -sum(kv[1][c] for _, c in stalls)):
    s = sum(v[c] for _, c in stalls)
    if s < 0.005 * tot: continue
    print(f"{f:24s} {v['inst']:10d} {s/tot:8.3f}  " + " ".join(f"{v[c]/tot:8.3f}" for c in cols))
