#!/bin/bash
# round 2, call 15 (2 GPUs): the one-shot NVLink all-reduce: parity test + weak-scaling bench beside NCCL
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_round2_gpu.py -m gpu -q -k two_nccl --timeout 600 > gpurun_out/r2c15_pytest_g2.log 2>&1
tail -15 gpurun_out/r2c15_pytest_g2.log
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29511 bench.py --gpus 2 $B > gpurun_out/r2c15_bench_c2_g2.json 2> gpurun_out/r2c15_bench_c2_g2.err
PINNJET_ALLREDUCE=nccl timeout 300 $R --master-port 29512 bench.py --gpus 2 $B > gpurun_out/r2c15_bench_c2_g2_nccl.json 2> gpurun_out/r2c15_bench_c2_g2_nccl.err
timeout 200 python bench.py $B > gpurun_out/r2c15_bench_c2_g1.json 2> gpurun_out/r2c15_bench_c2_g1.err
python - <<'PY'
import json
for w in ("c2_g1","c2_g2","c2_g2_nccl"):
    f=f"gpurun_out/r2c15_bench_{w}.json"
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(w, "value %.4g ms/step %.4f e2e %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("collective"))
    except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
