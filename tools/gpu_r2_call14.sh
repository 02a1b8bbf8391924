#!/bin/bash
# round 2, call 14: tensor-core kernels as the default: full suite, traces, bench, ncu captures (C2)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/r2c14_pytest_default.log 2>&1
tail -5 gpurun_out/r2c14_pytest_default.log
for k in k1 k2; do timeout 300 python tools/gpu_trace_tc.py c2 $k > gpurun_out/r2c14_trace_${k}_c2.txt 2>&1; done
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator"
for w in c2 c5 c4; do
  timeout 200 python bench.py $B --workload $w > gpurun_out/r2c14_bench_${w}.json 2> gpurun_out/r2c14_bench_${w}.err
done
python - <<'PY'
import json
for w in ("c2","c5","c4"):
    f=f"gpurun_out/r2c14_bench_{w}.json"
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(w, "ms/step %.4f e2e %.4f k1 %.1f us k2 %.1f us loss %r" % (d["ms_per_step"], d["e2e"]["ms_per_step"], r["launch_ms"]*1e3, r["k2"]["launch_ms"]*1e3, d.get("loss")))
    except Exception as e: print("ERR", f, e)
PY
for k in k1tc3_forward k2tc2_backward; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2c14_prof_$k -f \
        python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator > gpurun_out/r2c14_ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c14_launches_c2.csv \
        python bench.py --steps 5 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator > gpurun_out/r2c14_ncu_launches.log 2>&1
ls -la gpurun_out/r2c14* | head -20
