#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PINNJET_TC=1
for k in c2 c5; do
  echo "=== debug $k (TC)" >> gpurun_out/debug_tc.log
  timeout 120 python tools/gpu_debug.py $k 300 >> gpurun_out/debug_tc.log 2>&1
  echo "exit $?" >> gpurun_out/debug_tc.log
done
grep -v "^$" gpurun_out/debug_tc.log | cut -c1-180 | tail -60
