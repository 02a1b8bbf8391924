#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
p=29511
for w in c2 c3; do
  p=$((p+1))
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 100 --warmup 10 --workload $w > gpurun_out/bench_${w}_g2.json 2> gpurun_out/bench_${w}_g2.err
  grep -i "capture failed" gpurun_out/bench_${w}_g2.err | head -2
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${w}_g2.json").read().strip().splitlines()[-1])
print("$w", d["value"], d["ms_per_step"], d["e2e"]["value"], "graph", d["config"]["cuda_graph"])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-260
