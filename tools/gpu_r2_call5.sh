#!/bin/bash
# round 2, call 5: full suite on the tensor-core pair + ncu captures of both kernels (C2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== pytest PINNJET_TC=2 (whole GPU suite)"
PINNJET_TC=2 timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2c5_pytest_tc2.log 2>&1
tail -15 gpurun_out/r2c5_pytest_tc2.log
export PINNJET_TC=2
for k in k1tc3_forward k2tc2_backward; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2c5_prof_$k -f \
        python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator > gpurun_out/r2c5_ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c5_launches_c2.csv \
        python bench.py --steps 5 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator > gpurun_out/r2c5_ncu_launches.log 2>&1
ls -la gpurun_out/r2c5*
