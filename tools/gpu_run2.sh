#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/gpu_precision.py > gpurun_out/precision.log 2>&1
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
tail -c 3000 gpurun_out/bench_c2.json
timeout 600 python bench.py --steps 100 --warmup 10 --no-graph --cpu-seconds 1 > gpurun_out/bench_c2_nograph.json 2>> gpurun_out/bench_c2.err
for w in c3 c4 c5 c1; do
  timeout 600 python bench.py --steps 50 --warmup 5 --workload $w --cpu-seconds 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
# launch list (cold-cache, serialised) of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-graph --cpu-seconds 0.5 > gpurun_out/ncu_bench.log 2>&1
# full capture of K1 and K2 (3 launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k1_forward -s 3 -c 2 -o gpurun_out/prof_k1_c2 -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.5 > gpurun_out/ncu_k1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2_backward -s 3 -c 2 -o gpurun_out/prof_k2_c2 -f python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.5 > gpurun_out/ncu_k2.log 2>&1
ls -la gpurun_out
