#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2c3_pytest_all.log 2>&1
tail -30 gpurun_out/r2c3_pytest_all.log
