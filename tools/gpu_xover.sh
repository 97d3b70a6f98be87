#!/bin/bash
for S in 8192 12288 16384 24576; do
  for mode in 0 1; do
    timeout 300 python tools/fused_time.py cfg3 $S $mode 4 2>&1 | tail -1 | cut -c1-140
  done
done
