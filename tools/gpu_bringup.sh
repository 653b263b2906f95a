#!/bin/bash
# Bring-up script for a fresh B200 box: every stage is bounded by `timeout` and logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== gemm tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k gemm --no-header -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/t_gemm_all.log; tail -25 gpurun_out/t_gemm_all.log
if grep -q failed gpurun_out/t_gemm_all.log; then
echo "== MN-major hypotheses"
for v in "4096 512" "512 4096" "4096 1024" "128 512" "512 128"; do set -- $v; echo "-- MN_LBO=$1 MN_SBO=$2"; FLEXDM_MN_LBO=$1 FLEXDM_MN_SBO=$2 timeout 120 python -m pytest tests/test_gpu_parity.py -q -k "gemm and tcgen05" --no-header -p no:cacheprovider 2>&1 | tail -12; done > gpurun_out/t_mn.log 2>&1; cat gpurun_out/t_mn.log
fi
echo "== engine tests with tcgen05"; timeout 900 python -m pytest tests -q -m gpu -k "not gemm" --no-header -p no:cacheprovider > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/t_tc.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
