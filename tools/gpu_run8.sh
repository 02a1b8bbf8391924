#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/gpu_timing.py c2 c3 c5 > gpurun_out/timing.log 2>&1
grep "K2 CTA0" gpurun_out/timing.log | cut -c1-220
bash tools/gpu_run7.sh 2>&1 | head -5
