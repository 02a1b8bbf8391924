#!/bin/bash
# final evidence run (1 GPU): tests, bench all workloads, ncu launch list + full captures of K1/K2 for C2 and C3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_run7.sh 2>&1 | head -4
timeout 600 python bench.py --steps 100 --warmup 10 --workload c1 --cpu-seconds 2 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-graph --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
for w in c2 c3; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_forward -s 3 -c 1 -o gpurun_out/prof_k1_$w -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.5 --workload $w > gpurun_out/ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_backward -s 3 -c 1 -o gpurun_out/prof_k2_$w -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.5 --workload $w > gpurun_out/ncu_k2.log 2>&1
done
ls gpurun_out/*.ncu-rep
