#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_device_loop_gpu.py -m gpu -q --timeout 300 > gpurun_out/r2c16_pytest_devloop.log 2>&1
tail -30 gpurun_out/r2c16_pytest_devloop.log
timeout 400 python bench.py --steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 200 --no-gpu-comparator > gpurun_out/r2c16_bench_c2.json 2> gpurun_out/r2c16_bench_c2.err
tail -3 gpurun_out/r2c16_bench_c2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2c16_bench_c2.json"))
print("ms/step", d["ms_per_step"], "e2e", d["e2e"]["ms_per_step"])
print("fit", {k:v for k,v in d["fit"].items() if k!="device_loop"})
print("fit device_loop", d["fit"].get("device_loop"))
PY
