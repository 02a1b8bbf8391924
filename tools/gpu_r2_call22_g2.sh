#!/bin/bash
# round 2, call 22 (2 GPUs): the fused reduce + all-reduce kernel after the 32 x 8 restructuring: parity test + weak-scaling bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k pack_clears --timeout 200 > gpurun_out/r2c22_pytest_pack.log 2>&1
tail -4 gpurun_out/r2c22_pytest_pack.log
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "two_nccl" --timeout 500 > gpurun_out/r2c22_pytest_g2.log 2>&1
tail -6 gpurun_out/r2c22_pytest_g2.log
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29511 bench.py --gpus 2 $B > gpurun_out/r2c22_bench_c2_g2.json 2> gpurun_out/r2c22_bench_c2_g2.err
python - <<'PY'
import json
for w in ("c2_g2",):
    f=f"gpurun_out/r2c22_bench_{w}.json"
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(w, "value %.4g ms/step %.4f e2e %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("step_ms_stats"), d.get("strong_scaling"))
    except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
