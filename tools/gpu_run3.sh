#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for w in c2 c3 c4 c5; do
  timeout 600 python bench.py --steps 100 --warmup 10 --workload $w --cpu-seconds 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json")); r=d["roofline"]
print("$w pts/s %.3e ms/step %.4f e2e %.3e k1 %.1f us (%.1f%% fp32) k2 %.1f us (%.1f%%) cpu %.0f (%d thr)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["launch_ms"]*1e3,100*r["frac_of_fp32_ffma_peak"],r["k2"]["launch_ms"]*1e3,100*r["k2"]["frac_of_fp32_ffma_peak"],d["cpu_baseline"]["value"],d["cpu_baseline"]["cores"]))
PY
done
