#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jit_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2c17_pytest_jit.log 2>&1
tail -25 gpurun_out/r2c17_pytest_jit.log
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
for w in c2 c5 c4; do
  timeout 200 python bench.py $B --workload $w > gpurun_out/r2c17_bench_${w}.json 2> gpurun_out/r2c17_bench_${w}.err
  PINNJET_JIT=0 timeout 200 python bench.py $B --workload $w > gpurun_out/r2c17_bench_${w}_nojit.json 2> gpurun_out/r2c17_bench_${w}_nojit.err
done
python - <<'PY'
import json
for w in ("c2","c2_nojit","c5","c5_nojit","c4","c4_nojit"):
    f=f"gpurun_out/r2c17_bench_{w}.json"
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(w, "ms/step %.4f e2e %.4f k1 %.1f us (%.3f fp32) k2 %.1f us jit %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], r["launch_ms"]*1e3, r["frac_of_fp32_ffma_peak"], r["k2"]["launch_ms"]*1e3, d["specialised_forward_kernel"]))
    except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-800:])
PY
