#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PINNJET_TC=1 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_tc.log 2>&1
tail -5 gpurun_out/pytest_gpu_tc.log
for w in c2 c5; do
  for tc in 0 1; do
  PINNJET_TC=$tc timeout 600 python bench.py --steps 100 --warmup 10 --workload $w --cpu-seconds 1 > gpurun_out/bench_${w}_tc$tc.json 2> gpurun_out/bench_${w}_tc$tc.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${w}_tc$tc.json")); r=d["roofline"]
    print("$w TC=$tc pts/s %.3e ms/step %.4f e2e %.3e k1 %.1f us (%.1f%% fp32) k2 %.1f us loss %.6e"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["launch_ms"]*1e3,100*r["frac_of_fp32_ffma_peak"],r["k2"]["launch_ms"]*1e3, d["loss"]))
except Exception as e:
    print("ERR $w $tc", e); print(open("gpurun_out/bench_${w}_tc$tc.err").read()[-600:])
PY
  done
done
