#!/bin/bash
for cfg in cfg5 cfg2; do
for S in 4096 8192 16384; do
  echo "$cfg S=$S fused:"; BORE_LB_FUSED_MAX=1000000 timeout 300 python tools/fused_time.py $cfg $S 2 4 2>&1 | tail -1 | cut -c1-72
  echo "$cfg S=$S rounds+handover:"; BORE_LB_FUSED_MAX=1024 timeout 300 python tools/fused_time.py $cfg $S 2 4 2>&1 | tail -1 | cut -c1-72
done
done
