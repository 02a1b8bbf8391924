#!/bin/bash
# round 2, call 18 (2 GPUs): K2b + collective as one kernel (pj_backward_allreduce): parity test + weak-scaling bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_round2_gpu.py -m gpu -q -k "two_nccl and c2" --timeout 500 > gpurun_out/r2c18_pytest_g2.log 2>&1
tail -12 gpurun_out/r2c18_pytest_g2.log
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29511 bench.py --gpus 2 $B > gpurun_out/r2c18_bench_c2_g2.json 2> gpurun_out/r2c18_bench_c2_g2.err
PINNJET_FUSED_AR=0 PINNJET_BENCH_ALIGN=0 timeout 300 $R --master-port 29512 bench.py --gpus 2 $B --no-strong > gpurun_out/r2c18_bench_c2_g2_unfused.json 2> gpurun_out/r2c18_bench_c2_g2_unfused.err
timeout 200 python bench.py $B --no-strong > gpurun_out/r2c18_bench_c2_g1.json 2> gpurun_out/r2c18_bench_c2_g1.err
python - <<'PY'
import json
for w in ("c2_g1","c2_g2","c2_g2_unfused"):
    f=f"gpurun_out/r2c18_bench_{w}.json"
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(w, "value %.4g ms/step %.4f e2e %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d.get("step_ms_stats"), d.get("collective"), d.get("strong_scaling"))
    except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
