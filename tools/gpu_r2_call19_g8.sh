#!/bin/bash
# round 2, call 19 (8 GPUs): the data-parallel path on every GPU of the box: 8-rank parity test + weak/strong scaling bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "two_nccl and c2" --timeout 350 > gpurun_out/r2c19_pytest_g8.log 2>&1
tail -8 gpurun_out/r2c19_pytest_g8.log
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29511 bench.py --gpus 8 $B > gpurun_out/r2c19_bench_c2_g8.json 2> gpurun_out/r2c19_bench_c2_g8.err
python - <<'PY'
import json
for w in ("c2_g8",):
    f=f"gpurun_out/r2c19_bench_{w}.json"
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(w, "value %.4g ms/step %.4f e2e %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("step_ms_stats"), d.get("collective"), d.get("strong_scaling"))
    except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-2500:])
PY
