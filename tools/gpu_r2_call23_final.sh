#!/bin/bash
# round 2, call 23 (final, driver-like): full GPU suite with the specialised kernel as the solvers' default, the default bench
# invocation (fit legs, comparator, strong-scaling legs, cpu baseline), smoke()
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/r2c23_pytest_gpu.log 2>&1
tail -12 gpurun_out/r2c23_pytest_gpu.log
timeout 240 python bench.py --gpus 1 --steps 100 --warmup 10 --cpu-seconds 4 > gpurun_out/r2c23_bench_default.json 2> gpurun_out/r2c23_bench_default.err
python - <<'PY'
import json
f="gpurun_out/r2c23_bench_default.json"
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
    print("default bench: value %.4g ms/step %.4f e2e %.4f k1 %.1f us (%.3f fp32) k2 %.1f us loss %.6g launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r["launch_ms"]*1e3, r["frac_of_fp32_ffma_peak"], r["k2"]["launch_ms"]*1e3, d["loss"], d["gpu_launches"]))
    print("fit", json.dumps(d.get("fit"))[:900])
    print("cmp", json.dumps(d.get("gpu_autograd_baseline"))[:300])
    print("cpu", json.dumps(d.get("cpu_baseline"))[:400])
    print("strong", json.dumps(d.get("strong_scaling"))[:600])
except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
