#!/bin/bash
# round 2, call 2: K2-TC timing + ncu captures of the shipped default kernels (profiles/r02)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="--steps 50 --warmup 5 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
echo "=== k2tc bench c2 / c5 (TC=2 forward too)"
( export PINNJET_LIB=$PWD/neurodiffeq_b200/csrc/libpinnjet_exp.so PINNJET_TC_BWD=1
  timeout 200 python bench.py $B > gpurun_out/r2c2_bench_c2_k2tc.json 2> gpurun_out/r2c2_bench_c2_k2tc.err
  timeout 200 python bench.py $B --workload c5 > gpurun_out/r2c2_bench_c5_k2tc.json 2> gpurun_out/r2c2_bench_c5_k2tc.err
  PINNJET_TC=2 timeout 200 python bench.py $B > gpurun_out/r2c2_bench_c2_k2tc_tc2.json 2> gpurun_out/r2c2_bench_c2_k2tc_tc2.err
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k2tc_backward -s 2 -c 1 -o gpurun_out/r2c2_prof_k2tc -f \
        python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator > gpurun_out/r2c2_ncu_k2tc.log 2>&1
)
python - <<'PY'
import json
for f in ("c2_k2tc","c5_k2tc","c2_k2tc_tc2"):
    try:
        d=json.load(open(f"gpurun_out/r2c2_bench_{f}.json")); r=d["roofline"]
        print(f, "ms/step %.4f k1 %.1f us k2 %.1f us" % (d["ms_per_step"], r["launch_ms"]*1e3, r["k2"]["launch_ms"]*1e3))
    except Exception as e: print("ERR", f, e)
PY
echo "=== ncu shipped kernels"
for k in k1_forward k2_backward; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2c2_prof_$k -f \
        python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator > gpurun_out/r2c2_ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c2_launches_c2.csv \
        python bench.py --steps 5 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator > gpurun_out/r2c2_ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
