#!/usr/bin/env python
"""Executed warp instructions of one ncu capture grouped by SASS opcode.
usage: ncu_by_opcode.py <report.ncu-rep> [top]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = [i for i, r in enumerate(rows) if "Address" in r][0]
hdr = rows[h]
isrc, ie = hdr.index("Source"), hdr.index("Instructions Executed")
agg = collections.Counter()
tot = 0
for r in rows[h + 1:]:
    if len(r) <= ie: continue
    toks = r[isrc].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:2])
    n = int(r[ie]); agg[op] += n; tot += n
print("total", tot)
for op, n in agg.most_common(top):
    print(f"{op:24s} {n:12d} {n/tot:6.3f}")
