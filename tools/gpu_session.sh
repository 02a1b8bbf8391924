#!/bin/bash
# Recipes for one `gpurun` call on a B200 box (everything written under gpurun_out/, which gpurun merges back).
#   gpurun --timeout 600 -- 'bash tools/gpu_session.sh <mode> [args]'
# modes
#   tests            pytest -m gpu (all) + smoke
#   tests-tc N       the parity suite with the tensor-core forward kernel variant N (PINNJET_TC=N)
#   tests-new        the GPU tests of features that were only CPU-verified so far (PINNJET_TEST_UNVALIDATED=1)
#   bench [wl ...]   bench.py for the given workloads (default: c2 c3 c4 c5 c1), one JSON per workload + a summary line
#   launches         ncu launch list of bench.py (gpu__time_duration per kernel; cold, serialised: shares only)
#   ncu KERNEL [TC]  one `ncu --set full` capture of a kernel matching regex KERNEL (k1_forward, k2_backward, k1tc2_forward)
#   timing [wl ...]  phase cycle counters of the PJ_TIMING diagnostic build (tools/gpu_timing.py)
#   precision        fp32 kernels vs fp64 oracle vs the oracle's own fp32 run (tools/gpu_precision.py)
#   scale2           2-GPU weak scaling of bench.py (needs gpurun --gpus 2)
#   probes           the stand-alone tcgen05 probes under experiments/tcgen05_probe
#   k2tc             bring-up of the experimental tensor-core reverse kernel: build `python neurodiffeq_b200/csrc/build.py
#                    --experimental` HERE first (libpinnjet_exp.so travels with the snapshot), then parity suite + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
mode=${1:-tests}; shift

summary() {   # one line per bench JSON
python - "$@" <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.load(open(f)); r = d["roofline"]
        print("%s pts/s %.3e ms/step %.4f e2e %.3e k1 %.1f us (%.1f%% fp32) k2 %.1f us (%.1f%%) cpu %s" % (
            f, d["value"], d["ms_per_step"], d["e2e"]["value"], r["launch_ms"] * 1e3, 100 * r["frac_of_fp32_ffma_peak"],
            r["k2"]["launch_ms"] * 1e3, 100 * r["k2"]["frac_of_fp32_ffma_peak"], (d.get("cpu_baseline") or {}).get("value")))
    except Exception as e:
        print("ERR", f, e)
PY
}

case "$mode" in
  tests)
    timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
    timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1 ;;
  tests-tc)
    PINNJET_TC=${1:-2} timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_solvers_gpu.py tests/test_properties_gpu.py \
        -m gpu -q > gpurun_out/pytest_gpu_tc.log 2>&1; tail -5 gpurun_out/pytest_gpu_tc.log ;;
  tests-new)
    PINNJET_TEST_UNVALIDATED=1 timeout 600 python -m pytest tests/test_conditions_gpu.py -m gpu -q > gpurun_out/pytest_gpu_new.log 2>&1
    tail -12 gpurun_out/pytest_gpu_new.log ;;
  bench)
    wls=${@:-c2 c3 c4 c5 c1}; files=""
    for w in $wls; do
      timeout 600 python bench.py --steps 100 --warmup 10 --workload $w --cpu-seconds 2 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
      files="$files gpurun_out/bench_$w.json"
    done
    summary $files ;;
  launches)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c2.csv \
        python bench.py --steps 5 --warmup 3 --no-graph --cpu-seconds 0.5 --fit-epochs 0 --no-gpu-comparator > gpurun_out/ncu_bench.log 2>&1
    tail -3 gpurun_out/ncu_bench.log ;;
  ncu)
    k=${1:-k1_forward}; tc=${2:-0}
    PINNJET_TC=$tc timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k -f \
        python bench.py --steps 3 --warmup 3 --no-graph --cpu-seconds 0.3 --fit-epochs 0 --no-gpu-comparator > gpurun_out/ncu_$k.log 2>&1
    ls -la gpurun_out/prof_$k.ncu-rep ;;
  timing)
    timeout 600 python tools/gpu_timing.py ${@:-c2 c3 c4 c5} > gpurun_out/timing.log 2>&1; tail -20 gpurun_out/timing.log ;;
  precision)
    timeout 600 python tools/gpu_precision.py > gpurun_out/precision.log 2>&1; tail -20 gpurun_out/precision.log ;;
  scale2)
    p=29511
    for w in c2 c3; do
      p=$((p+1))
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p \
          bench.py --gpus 2 --steps 100 --warmup 10 --workload $w > gpurun_out/bench_${w}_g2.json 2> gpurun_out/bench_${w}_g2.err
    done
    summary gpurun_out/bench_c2_g2.json gpurun_out/bench_c3_g2.json ;;
  probes)
    bash experiments/tcgen05_probe/run.sh; bash experiments/tcgen05_probe/run_wgrad.sh ;;
  k2tc)
    export PINNJET_LIB=$PWD/neurodiffeq_b200/csrc/libpinnjet_exp.so PINNJET_TC_BWD=1
    ls -la $PINNJET_LIB || exit 2
    timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "c2 or c5" > gpurun_out/pytest_gpu_k2tc.log 2>&1
    tail -15 gpurun_out/pytest_gpu_k2tc.log
    timeout 200 python bench.py --steps 50 --warmup 5 --cpu-seconds 0.5 --fit-epochs 0 --no-gpu-comparator \
        > gpurun_out/bench_c2_k2tc.json 2> gpurun_out/bench_c2_k2tc.err
    summary gpurun_out/bench_c2_k2tc.json ;;
  *) echo "unknown mode $mode"; exit 2 ;;
esac
