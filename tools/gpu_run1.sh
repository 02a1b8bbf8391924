#!/bin/bash
# first GPU contact: stage-by-stage debug + parity tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for k in c2 c1 c5 c3 c4; do
  echo "=== debug $k" >> gpurun_out/debug.log
  timeout 300 python tools/gpu_debug.py $k 300 >> gpurun_out/debug.log 2>&1
  echo "exit $?" >> gpurun_out/debug.log
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
