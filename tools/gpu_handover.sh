#!/bin/bash
mkdir -p gpurun_out
for h in 2048 4096 6144 8192 12288; do
BORE_LB_HANDOVER=$h timeout 300 python bench.py --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_h$h.json 2> gpurun_out/bench_h$h.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_h$h.json').read().strip().splitlines()[-1])
print($h, round(d['ms_per_step'],2), d['phases']['lbfgsb_rounds'], [(k['name'][:12],round(k['ms_per_step'],2)) for k in d['kernels']], d['phases']['evals_per_step_per_gpu'])
PY
done
