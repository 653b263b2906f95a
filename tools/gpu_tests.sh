#!/bin/bash
# parity tests only (fast): usage gpu_tests.sh [pytest -k expression]
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -rxXw -m gpu --no-header -p no:cacheprovider ${1:+-k "$1"} > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_gpu.log; grep -E "^FAILED|^XPASS|^XFAIL|^E  |UserWarning: TF32 operand rounding" gpurun_out/t_gpu.log | cut -c1-300 | head -30
