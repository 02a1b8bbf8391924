#!/bin/bash
# one ncu capture of the transposed-epilogue tensor-core forward kernel (C2), for the round-2 work list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PINNJET_TC=2 timeout 170 ncu --set full --clock-control none --import-source on -k regex:k1tc2_forward -s 2 -c 1 \
    -o gpurun_out/prof_k1tc2_c2 -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.3 --fit-epochs 0 \
    --no-gpu-comparator > gpurun_out/ncu_k1tc2.log 2>&1
ls -la gpurun_out/prof_k1tc2_c2.ncu-rep
