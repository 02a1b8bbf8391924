#!/bin/bash
# last GPU call of round 1 (about 5 GPU-minutes left): default path first, then the opt-in tensor-core variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
t0=$SECONDS
step() { echo "== $1 (t=$((SECONDS-t0))s)"; }

step "validated suite, default path"
timeout 150 python -m pytest tests/test_kernels_gpu.py tests/test_solvers_gpu.py tests/test_properties_gpu.py -m gpu -q -x \
    > gpurun_out/pytest_gpu_core_r1b.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/pytest_gpu_core_r1b.log
step "new tests: Neumann conditions, named losses"
timeout 120 python -m pytest tests/test_conditions_gpu.py tests/test_losses_gpu.py -m gpu -q \
    > gpurun_out/pytest_gpu_new_r1b.log 2>&1; echo "rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu_new_r1b.log | tail -12
step "tensor-core variant 2: parity suite"
PINNJET_TC=2 timeout 100 python -m pytest tests/test_kernels_gpu.py tests/test_solvers_gpu.py tests/test_properties_gpu.py -m gpu -q \
    > gpurun_out/pytest_gpu_tc2.log 2>&1; echo "rc=$?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu_tc2.log | tail -8
step "bench default (with fit leg and torch-CUDA comparator)"
timeout 150 python bench.py --cpu-seconds 6 > gpurun_out/bench_c2_r1b.json 2> gpurun_out/bench_c2_r1b.err; echo "rc=$?"
step "bench TC=2 c2"
PINNJET_TC=2 timeout 60 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.5 --fit-epochs 0 --no-gpu-comparator \
    > gpurun_out/bench_c2_tc2.json 2> gpurun_out/bench_c2_tc2.err; echo "rc=$?"
step "bench TC=2 c5"
PINNJET_TC=2 timeout 60 python bench.py --workload c5 --steps 50 --warmup 5 --cpu-seconds 0.5 --fit-epochs 0 --no-gpu-comparator \
    > gpurun_out/bench_c5_tc2.json 2> gpurun_out/bench_c5_tc2.err; echo "rc=$?"
step "done"
for f in bench_c2_r1b bench_c2_tc2 bench_c5_tc2; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/$f.json")); r = d["roofline"]
    print("$f pts/s %.3e ms/step %.4f e2e %.3e k1 %.1f us k2 %.1f us" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_ms"] * 1e3, r["k2"]["launch_ms"] * 1e3))
    if d.get("fit"): print("  fit", d["fit"])
    if d.get("gpu_autograd_baseline"): print("  gpu autograd", d["gpu_autograd_baseline"])
except Exception as e:
    print("ERR $f", e)
PY
done
