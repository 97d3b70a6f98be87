#!/bin/bash
# one gpurun call: L-BFGS-B parity tests, bench, launch list, one ncu capture of the stepper
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lbfgsb.py tests/test_gpu_surface.py tests/test_gpu_batched.py -m gpu -q -x > gpurun_out/pytest_lb.log 2>&1; tail -4 gpurun_out/pytest_lb.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','phases')}); print([ (k['name'][:20],round(k['ms_per_step'],2),round(k['frac'],4)) for k in d['kernels']]); print(d['e2e'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv python tools/profile_target.py cfg3 > gpurun_out/launch_run.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lbfgsb_warp -s 25 -c 1 -o gpurun_out/prof_step_r25 -f python tools/profile_target.py cfg3 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
