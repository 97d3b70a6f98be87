#!/bin/bash
mkdir -p gpurun_out
for np in -1 0 1; do
  BORE_LB_NPHASE=$np timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_np$np.json 2> gpurun_out/bench_np$np.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_np$np.json').read().strip().splitlines()[-1])
print($np, round(d['ms_per_step'],1), [(k['name'][:12],round(k['ms_per_step'],2)) for k in d['kernels']], d['phases']['evals_per_step_per_gpu'])
PY
done
