#!/bin/bash
# final evidence run of the round (default product path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1
tail -3 gpurun_out/pytest_gpu_final.log
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_c2_final.json 2> gpurun_out/bench_c2_final.err
for w in c3 c4 c5 c1; do
  timeout 300 python bench.py --steps 50 --warmup 5 --workload $w --cpu-seconds 1 > gpurun_out/bench_${w}_final.json 2> gpurun_out/bench_${w}_final.err
done
for w in c2 c3 c4 c5 c1; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${w}_final.json")); r=d["roofline"]
    print("$w pts/s %.3e ms/step %.4f e2e %.3e k1 %.1f us (%.1f%%) k2 %.1f us (%.1f%%) cpu %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["launch_ms"]*1e3,100*r["frac_of_fp32_ffma_peak"],r["k2"]["launch_ms"]*1e3,100*r["k2"]["frac_of_fp32_ffma_peak"], (d["cpu_baseline"] or {}).get("value")))
except Exception as e: print("ERR $w", e)
PY
done
