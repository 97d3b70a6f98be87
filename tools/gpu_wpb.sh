#!/bin/bash
mkdir -p gpurun_out
for w in 11 5 3 2; do
  BORE_LB_WPB=$w timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_w$w.json').read().strip().splitlines()[-1])
print($w, round(d['ms_per_step'],1), [(k['name'][:12],round(k['ms_per_step'],2)) for k in d['kernels']], d['phases']['evals_per_step_per_gpu'])
PY
done
