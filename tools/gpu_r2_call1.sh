#!/bin/bash
# round 2, call 1: everything un-gated, K2-TC bring-up, sanitizer, baseline bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2c1_smi.txt 2>&1
PINNJET_TEST_UNVALIDATED=1 timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c1_pytest_all.log 2>&1
tail -25 gpurun_out/r2c1_pytest_all.log
echo "=== k2tc"
( export PINNJET_LIB=$PWD/neurodiffeq_b200/csrc/libpinnjet_exp.so PINNJET_TC_BWD=1
  timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "golden and (c2 or c5)" > gpurun_out/r2c1_k2tc.log 2>&1
  tail -15 gpurun_out/r2c1_k2tc.log )
echo "=== sanitizer"
timeout 400 compute-sanitizer --tool memcheck --log-file gpurun_out/r2c1_memcheck.log python tools/sanitize_case.py c2 c5 c3 1500 > gpurun_out/r2c1_memcheck.out 2>&1
tail -3 gpurun_out/r2c1_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --log-file gpurun_out/r2c1_racecheck.log python tools/sanitize_case.py c2 c5 600 > gpurun_out/r2c1_racecheck.out 2>&1
tail -3 gpurun_out/r2c1_racecheck.log
echo "=== bench"
timeout 300 python bench.py --steps 100 --warmup 10 --cpu-seconds 1 --fit-epochs 100 --no-gpu-comparator > gpurun_out/r2c1_bench_c2.json 2> gpurun_out/r2c1_bench_c2.err
cat gpurun_out/r2c1_bench_c2.json | head -c 1500
