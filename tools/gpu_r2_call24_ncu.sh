#!/bin/bash
# round 2, call 24: ncu launch list + full capture of the specialised forward kernel for the FINAL build (4 launches per step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
A="--steps 5 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator --no-strong"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c24_launches_c2.csv \
        python bench.py $A > gpurun_out/r2c24_ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pj_k1_jit -s 2 -c 1 -o gpurun_out/r2c24_prof_pj_k1_jit -f \
        python bench.py $A > gpurun_out/r2c24_ncu_k1.log 2>&1
ls -la gpurun_out/r2c24*
