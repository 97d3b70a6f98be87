#!/bin/bash
# FFMA2 (packed fp32 FMA): rate microbenchmark, K2 parity, bench A/B with BORE_K2_FFMA2=0/1
mkdir -p gpurun_out
./tools/microbench/ffma2_rate > gpurun_out/ffma2_rate.txt 2>&1; cat gpurun_out/ffma2_rate.txt
timeout 600 python -m pytest tests/test_gpu_mlp_eval.py tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -2
for v in 0 1; do
BORE_K2_FFMA2=$v timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_f2_$v.json 2> gpurun_out/bench_f2_$v.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_f2_$v.json').read().strip().splitlines()[-1])
print('FFMA2=$v', round(d['ms_per_step'],1), [(k['name'][:12],round(k['ms_per_step'],2)) for k in d['kernels']], d['phases']['evals_per_step_per_gpu'], round(d['e2e']['ms_per_step'],1))
PY
done
