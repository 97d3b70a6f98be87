#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lbfgsb.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],1), [(k['name'][:12],round(k['ms_per_step'],2)) for k in d['kernels']], d['phases']['evals_per_step_per_gpu'], round(d['e2e']['ms_per_step'],1))
PY
