#!/bin/bash
mkdir -p gpurun_out
for t in 256 128 64 32; do
  BORE_FIT_THREADS=$t timeout 300 python bench.py --workload cfg4 --steps 2 --warmup 1 > gpurun_out/bench_cfg4_t$t.json 2> gpurun_out/bench_cfg4_t$t.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cfg4_t$t.json').read().strip().splitlines()[-1])
print($t, round(d['ms_per_step'],1), round(d['value'],1))
PY
done
