#!/bin/bash
# run on the GPU box: bash experiments/tcgen05_probe/run_wgrad.sh   (MN-major operands / M=64 accumulator layout)
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
{
for args in "0 64 64 2048" "1 64 64 2048" "64 0 64 2048" "64 64 64 2048" "2048 64 64 2048" "64 2048 64 2048" "1 64 64 32" \
            "2048 64 128 2048" "64 2048 128 2048" "2048 64 128 32"; do
  timeout 20 ./probe_wgrad $args; echo "   exit=$? args=[$args]"
done
} > ../../gpurun_out/tcgen05_probe_wgrad.log 2>&1
cat ../../gpurun_out/tcgen05_probe_wgrad.log
