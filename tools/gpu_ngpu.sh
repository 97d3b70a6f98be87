#!/bin/bash
# bench.py on N GPUs of one box (weak line + strong / cfg 4 sub-records): bash tools/gpu_ngpu.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02c_bench_cfg3_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02c_bench_cfg3_${N}gpu.json').read().strip().splitlines()[-1])
    print('N=$N', round(d['value']/1e6,1),'M evals/s', round(d['ms_per_step'],1),'ms', 'e2e', round(d['e2e']['value']/1e6,1), d['phases'])
    s=d.get('strong'); print('strong', s and (round(s['ms_per_step'],1), s['phases']))
    c=d.get('cfg4'); print('cfg4', c and (round(c['ms_per_step'],1), round(c['value']), c['phases_ms_rank0']))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_${N}gpu.err').read()[-2000:])
PY
