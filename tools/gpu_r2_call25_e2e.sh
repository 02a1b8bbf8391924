#!/bin/bash
# round 2, call 25 (last GPU seconds): the e2e leg with the gradient clear folded into the replayed graph
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 100 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.2 --fit-epochs 0 --no-gpu-comparator --no-strong > gpurun_out/r2c25_bench_c2.json 2> gpurun_out/r2c25_bench_c2.err
python - <<'PY'
import json
f="gpurun_out/r2c25_bench_c2.json"
try:
    d=json.loads([l for l in open(f) if l.startswith("{")][-1])
    print("ms/step %.4f e2e %.4f (%.4g points/s) loss %.6g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"], d["loss"]))
except Exception as e: print("ERR", e)
print(open(f.replace(".json",".err")).read()[-800:])
PY
