#!/bin/bash
BENCH_DEBUG=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_e2e.json 2> gpurun_out/bench.err; grep "step" gpurun_out/bench.err | head -30
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_e2e.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['warmup'], d['strong']['ms_per_step'], d['e2e']['ms_per_step'])
PY
