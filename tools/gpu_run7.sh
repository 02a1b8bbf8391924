#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in c2 c3 c4 c5; do
  timeout 600 python bench.py --steps 100 --warmup 10 --workload $w --cpu-seconds 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_$w.json")); r=d["roofline"]
print("$w pts/s %.3e ms/step %.4f e2e %.3e k1 %.1f us (%.1f%% fp32) k2 %.1f us (%.1f%%)"%(d["value"],d["ms_per_step"],d["e2e"]["value"],r["launch_ms"]*1e3,100*r["frac_of_fp32_ffma_peak"],r["k2"]["launch_ms"]*1e3,100*r["k2"]["frac_of_fp32_ffma_peak"]))
PY
done
PINNJET_BENCH_DIRTY_FLUSH=1 timeout 600 python bench.py --steps 100 --warmup 10 --workload c2 --cpu-seconds 1 > gpurun_out/bench_c2_dirty.json 2>/dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_c2_dirty.json')); r=d['roofline']
print('c2 dirty-flush: ms/step %.4f k1 %.1f k2 %.1f'%(d['ms_per_step'], r['launch_ms']*1e3, r['k2']['launch_ms']*1e3))"
