#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
for w in c2 c4; do
  for nc in 0 1; do
  PINNJET_NO_COMBINE=$nc timeout 600 python bench.py --steps 100 --warmup 10 --workload $w --cpu-seconds 1 > gpurun_out/bench_${w}_nc$nc.json 2> gpurun_out/bench_${w}_nc$nc.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${w}_nc$nc.json")); r=d["roofline"]
    print("$w no_combine=$nc pts/s %.3e ms/step %.4f k1 %.1f us (%.1f%% fp32) k2 %.1f us (%.1f%%) loss %.6e"%(d["value"],d["ms_per_step"],r["launch_ms"]*1e3,100*r["frac_of_fp32_ffma_peak"],r["k2"]["launch_ms"]*1e3,100*r["k2"]["frac_of_fp32_ffma_peak"], d["loss"]))
except Exception as e:
    print("ERR $w $nc", e); print(open("gpurun_out/bench_${w}_nc$nc.err").read()[-800:])
PY
  done
done
