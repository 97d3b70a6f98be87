#!/usr/bin/env python
"""Summarise an ncu launch list (gpu__time_duration.sum per launch) by kernel."""
import collections, csv, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); seq = []
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:44]
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg[name][0] += 1; agg[name][1] += v; seq.append((name, v))
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:46s} n={v[0]:4d} total={v[1]/1e3:9.3f} ms share={v[1]/tot:.3f}")
for key in sys.argv[2:]:
    print(key, [round(x) for n, x in seq if key in n][:70])
