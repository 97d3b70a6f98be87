#!/usr/bin/env python
"""Static SASS size per source function of one kernel (line info + function map of lbfgsb_core.h).
usage: sass_static.py <object-or-so> <kernel-substring> [cubin-substring]"""
import collections, os, re, subprocess, sys, tempfile
obj, kern_key = sys.argv[1:3]
ckey = sys.argv[3] if len(sys.argv) > 3 else ""
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
hdr_path = os.path.join(root, "bore_b200/csrc/lbfgsb_core.h")
src = open(hdr_path).read().split("\n")
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
for cubin in sorted(os.listdir(tmp)):
    if ckey not in cubin: continue
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(sass) if l.startswith(".text.") and kern_key in l]
    for start in starts:
        print(sass[start][:140])
        cur = None
        cnt = collections.Counter()
        ops = collections.Counter()
        for l in sass[start + 1:]:
            if l.startswith("//-----"): break
            m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
            if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                cnt[fn(cur)] += 1
                op = m.group(2).split()
                op = op[1] if op[0].startswith("@") else op[0]
                ops[op.split(".")[0]] += 1
        tot = sum(cnt.values())
        for k, v in cnt.most_common(30): print(f"  {k:28s} {v:7d} {16*v/1024:7.1f} KB")
        print(f"  total {tot} instr {16*tot/1024:.1f} KB; top ops:", ops.most_common(12))
