#!/bin/bash
# compute-sanitizer over tiny invocations of every kernel family; logs -> gpurun_out/sanitizer_*.log
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  for k in ${SAN_KERNELS:-k1 k1c k1u k1t k2 k3 k3f k3multi svgd data lstm}; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_target.py $k > gpurun_out/sanitizer_${tool}_$k.log 2>&1
    echo "$tool $k: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_$k.log | tail -1) $(grep -cE ' ok|loss' gpurun_out/sanitizer_${tool}_$k.log)"
  done
done
