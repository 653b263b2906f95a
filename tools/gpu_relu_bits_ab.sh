#!/bin/bash
mkdir -p gpurun_out
echo "== new test + gemm units"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_determinism.py -q -k "gemm or relu_gate" --no-header -p no:cacheprovider -x > gpurun_out/t_unit.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/t_unit.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|timeout|trap|FAILED" gpurun_out/t_unit.log | head -30; exit 0; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu.log; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | cut -c1-250 | head -20
tools/gpu_ab_env.sh FLEXDM_RELU_BITS=0
