#!/bin/bash
# round 2, call 20: programmatic dependent launch (K0 -> K1 -> finalize -> K2 -> K2b): full GPU suite with it on, bench A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2c20_pytest_gpu.log 2>&1
tail -15 gpurun_out/r2c20_pytest_gpu.log
B="--steps 100 --warmup 10 --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator --no-strong"
for w in c2 c5; do
  timeout 200 python bench.py $B --workload $w > gpurun_out/r2c20_bench_${w}.json 2> gpurun_out/r2c20_bench_${w}.err
  PINNJET_PDL=0 timeout 200 python bench.py $B --workload $w > gpurun_out/r2c20_bench_${w}_nopdl.json 2> gpurun_out/r2c20_bench_${w}_nopdl.err
done
timeout 200 python bench.py $B --workload c3 > gpurun_out/r2c20_bench_c3.json 2> gpurun_out/r2c20_bench_c3.err
python - <<'PY'
import json
for w in ("c2","c2_nopdl","c5","c5_nopdl","c3"):
    f=f"gpurun_out/r2c20_bench_{w}.json"
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1]); r=d["roofline"]
        print(w, "ms/step %.4f (median %.4f) e2e %.4f k1 %.1f us k2 %.1f us loss %.6g" % (d["ms_per_step"], d["step_ms_stats"]["median"], d["e2e"]["ms_per_step"], r["launch_ms"]*1e3, r["k2"]["launch_ms"]*1e3, d["loss"]))
    except Exception as e: print("ERR", f, e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
for k in pj_k1_jit k2tc2_backward; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2c20_prof_$k -f \
        python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator --no-strong > gpurun_out/r2c20_ncu_$k.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c20_launches_c2.csv \
        python bench.py --steps 5 --warmup 3 --no-graph --cpu-seconds 0.1 --fit-epochs 0 --no-gpu-comparator --no-strong > gpurun_out/r2c20_ncu_launches.log 2>&1
ls -la gpurun_out/r2c20* | head -30
