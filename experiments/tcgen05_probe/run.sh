#!/bin/bash
# run on the GPU box: bash experiments/tcgen05_probe/run.sh
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
{
for args in "0 64 1 0" "1 64 1 0" "0 64 0 0" "64 64 1 0" "0 64 1 1" "1 64 1 1"; do
  timeout 30 ./probe $args; echo "   exit=$? args=[$args]"
done
} > ../../gpurun_out/tcgen05_probe.log 2>&1
cat ../../gpurun_out/tcgen05_probe.log
