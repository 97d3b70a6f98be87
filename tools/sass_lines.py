#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (code-size hunting).
usage: sass_lines.py <object-or-so> <kernel-substring> [source-file-for-text] [top]"""
import collections, os, re, subprocess, sys, tempfile
obj, key = sys.argv[1:3]
srcp = sys.argv[3] if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cnt, ops, total = collections.Counter(), collections.Counter(), 0
for cb in os.listdir(tmp):
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cb)], capture_output=True, text=True).stdout.split("\n")
    starts = [i for i, l in enumerate(sass) if l.startswith(".text.") and key in l]
    if not starts: continue
    cur = None
    for l in sass[starts[0] + 1:]:
        if l.startswith("//-----"): break
        m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
        if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            cnt[cur] += 1; total += 1
            t = m.group(2).split()
            ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += 1
    break
print("instructions", total)
src = open(srcp).read().split("\n") if srcp else []
byfile = collections.Counter()
for c, n in cnt.items(): byfile[c[0] if c else None] += n
print(byfile.most_common())
for c, n in cnt.most_common(top):
    t = src[c[1] - 1].strip()[:90] if (c and srcp and c[0] == os.path.basename(srcp)) else ""
    print(c, n, t)
print(ops.most_common(20))
