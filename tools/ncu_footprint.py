#!/usr/bin/env python
"""Per-function executed code footprint + dynamic instructions + no_inst stalls of one capture.
usage: ncu_footprint.py <report.ncu-rep> <object> <kernel-substring> <steps>"""
import collections, csv, io, os, re, subprocess, sys, tempfile
rep, obj, kern_key, steps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
src = open(os.path.join(root, "bore_b200/csrc/lbfgsb_core.h")).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"(?:LB_HD|LB_FN|LB_NI|__device__|template).*?\b(\w+)\(", l)
    if m and not l.startswith(" "): funcs.append((i, m.group(1)))
def fn(c):
    if c is None: return "none"
    f, l = c
    if f != "lbfgsb_core.h": return f
    name = "?"
    for i, nm in funcs:
        if i <= l: name = nm
    return name
off2fn = {}
for cubin in os.listdir(tmp):
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    st = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern_key in l]
    if not st: continue
    cur = None
    for l in sass[st[0] + 1:]:
        if l.startswith("//-----"): break
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m: off2fn[int(m.group(1), 16)] = (fn(cur), cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[h]
ia, ie = hdr.index("Address"), hdr.index("Instructions Executed")
ini = hdr.index("stall_no_inst")
stalls = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
base = int(rows[h + 1][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
for r in rows[h + 1:]:
    if len(r) <= ie: continue
    f, _ = off2fn.get(int(r[ia], 16) - base, ("none", None))
    n = int(r[ie])
    a = agg[f]
    a[0] += 1; a[1] += n > 0.001 * steps; a[2] += n; a[3] += int(r[ini]); a[4] += sum(int(r[i]) for i in stalls)
T = [sum(v[k] for v in agg.values()) for k in range(5)]
print(f"{'function':26s} {'static':>7s} {'exec':>6s} {'dyn/step':>9s} {'no_inst':>8s} {'samples':>8s}")
for f, v in sorted(agg.items(), key=lambda kv: -kv[1][2]):
    print(f"{f:26s} {v[0]:7d} {v[1]:6d} {v[2]/steps:9.1f} {v[3]/max(T[3],1):8.3f} {v[4]/T[4]:8.3f}")
print(f"{'TOTAL':26s} {T[0]:7d} {T[1]:6d} {T[2]/steps:9.1f}  no_inst share of samples {T[3]/T[4]:.3f}")
