#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/gpu_timing.py c2 c3 c4 c5 > gpurun_out/timing.log 2>&1
cat gpurun_out/timing.log | tail -20
bash tools/gpu_run3.sh
