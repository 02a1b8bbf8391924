#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T="tests/test_kernels_gpu.py tests/test_conditions_gpu.py tests/test_properties_gpu.py"
PINNJET_TC=2 timeout 600 python -m pytest $T -m gpu -q --timeout 120 > gpurun_out/r2c9_pytest_tc2.log 2>&1
tail -8 gpurun_out/r2c9_pytest_tc2.log
for k in k1 k2; do PINNJET_TC=2 timeout 300 python tools/gpu_trace_tc.py c2 $k > gpurun_out/r2c9_trace_${k}_c2.txt 2>&1; done
B="--steps 50 --warmup 5 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
for w in c2 c5 c4; do
    PINNJET_TC=2 timeout 200 python bench.py $B --workload $w > gpurun_out/r2c9_bench_${w}_tc2.json 2> gpurun_out/r2c9_bench_${w}_tc2.err
done
python - <<'PY'
import json
for w in ("c2","c5","c4"):
    f=f"gpurun_out/r2c9_bench_{w}_tc2.json"
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(w, "ms/step %.4f k1 %.1f us k2 %.1f us loss %r" % (d["ms_per_step"], r["launch_ms"]*1e3, r["k2"]["launch_ms"]*1e3, d.get("loss")))
    except Exception as e: print("ERR", f, e)
PY
