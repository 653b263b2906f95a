#!/bin/bash
# Quick GPU visit: GEMM tests first (bounded), then the rest of the parity suite and a short bench.
mkdir -p gpurun_out
echo "== gemm tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k gemm --no-header -p no:cacheprovider -x > gpurun_out/t_gemm.log 2>&1; rc=$?; echo "rc=$rc"; tail -15 gpurun_out/t_gemm.log
if [ $rc -ne 0 ]; then exit 0; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -k "not gemm" --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
