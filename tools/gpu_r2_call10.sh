#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2 3; do PINNJET_TC=2 timeout 300 python tools/gpu_grad_diff.py c2 c5 2>&1 | grep -v "rel err [0-9.]*e-0[78]"; done > gpurun_out/r2c10_graddiff.txt 2>&1
cat gpurun_out/r2c10_graddiff.txt
