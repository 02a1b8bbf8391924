#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PINNJET_TC=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1tc_forward -s 2 -c 1 -o gpurun_out/prof_k1tc_c2 -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.3 > gpurun_out/ncu_k1tc.log 2>&1
ls -la gpurun_out/prof_k1tc_c2.ncu-rep
