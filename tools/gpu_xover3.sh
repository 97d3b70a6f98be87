#!/bin/bash
for S in 4096 6144 8192 10240; do
  echo "S=$S fused:"; BORE_LB_FUSED_MAX=1000000 timeout 300 python tools/fused_time.py cfg3 $S 2 4 2>&1 | tail -1 | cut -c1-70
  echo "S=$S rounds+handover:"; BORE_LB_FUSED_MAX=1024 timeout 300 python tools/fused_time.py cfg3 $S 2 4 2>&1 | tail -1 | cut -c1-70
done
