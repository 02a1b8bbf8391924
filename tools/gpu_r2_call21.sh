#!/bin/bash
# round 2, call 21: loss finalisation inside K1, K0 split over 4 CTAs per layer (+ clears the gradient buffer), K2b 32 x 8:
# full GPU suite (PDL off = default) and the bench on C2 / C5 / C4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2c21_pytest_gpu.log 2>&1
tail -25 gpurun_out/r2c21_pytest_gpu.log
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator --no-strong"
for w in c2 c5 c4; do
  timeout 200 python bench.py $B --workload $w > gpurun_out/r2c21_bench_${w}.json 2> gpurun_out/r2c21_bench_${w}.err
done
python - <<'PY'
import json
for w in ("c2","c5","c4"):
    f=f"gpurun_out/r2c21_bench_{w}.json"
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
        print(w, "ms/step %.4f (median %.4f) e2e %.4f k1 %.1f us (%.3f fp32) k2 %.1f us loss %.6g" % (d["ms_per_step"], d["step_ms_stats"]["median"], d["e2e"]["ms_per_step"], r["launch_ms"]*1e3, r["frac_of_fp32_ffma_peak"], r["k2"]["launch_ms"]*1e3, d["loss"]))
    except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
