#!/bin/bash
# round 2, third part: all GPU parity tests, smoke, sanitizer over the unit-split fit, the default bench (with the
# CPU baseline and the LSTM record), launch list, full ncu capture of the unit-split fit
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
SAN_KERNELS="k1u" bash tools/gpu_sanitize.sh
timeout 900 python bench.py > gpurun_out/r02c_bench_cfg3.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench_cfg3.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','phases')}); print([ (k['name'][:20],round(k['ms_per_step'],2),round(k['frac'],4)) for k in d['kernels']]); print(d['e2e']); print(d.get('cpu_baseline')); print(d.get('lstm')); print(d.get('strong')); print(d.get('cfg4'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-3000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02c_launches_cfg3.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
tail -1 gpurun_out/launch_run.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fit_unit -c 1 -o gpurun_out/prof_fitu_r02c -f python tools/profile_target.py cfg3 2048 > gpurun_out/ncu_fitu.log 2>&1
tail -1 gpurun_out/ncu_fitu.log
