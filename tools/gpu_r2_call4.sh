#!/bin/bash
# round 2, call 4: first run of the K1-TC3 / K2-TC2 pair
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="tests/test_kernels_gpu.py tests/test_conditions_gpu.py tests/test_properties_gpu.py"
for lvl in 1 2; do
  echo "=== pytest PINNJET_TC=$lvl"
  PINNJET_TC=$lvl timeout 600 python -m pytest $T -m gpu -q --timeout 120 > gpurun_out/r2c4_pytest_tc$lvl.log 2>&1
  tail -25 gpurun_out/r2c4_pytest_tc$lvl.log
done
B="--steps 50 --warmup 5 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
for w in c2 c5 c4; do
  for lvl in 1 2; do
    PINNJET_TC=$lvl timeout 200 python bench.py $B --workload $w > gpurun_out/r2c4_bench_${w}_tc$lvl.json 2> gpurun_out/r2c4_bench_${w}_tc$lvl.err
  done
done
python - <<'PY'
import json
for w in ("c2","c5","c4"):
  for l in (1,2):
    f=f"gpurun_out/r2c4_bench_{w}_tc{l}.json"
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(w, l, "ms/step %.4f k1 %.1f us k2 %.1f us loss %r" % (d["ms_per_step"], r["launch_ms"]*1e3, r["k2"]["launch_ms"]*1e3, d.get("loss")))
    except Exception as e: print("ERR", f, e)
PY
