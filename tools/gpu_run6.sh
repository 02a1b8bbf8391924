#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/gpu_timing.py c2 c4 c5 > gpurun_out/timing.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_forward -s 3 -c 1 -o gpurun_out/prof_k1_c2 -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.5 > gpurun_out/ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_backward -s 3 -c 1 -o gpurun_out/prof_k2_c2 -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.5 > gpurun_out/ncu_k2.log 2>&1
for w in c2 c4; do
  timeout 600 python bench.py --steps 100 --warmup 10 --workload $w --cpu-seconds 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
