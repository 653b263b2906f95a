#!/bin/bash
# Quick GPU visit: kernel unit tests first (bounded), then the rest of the parity suite, a short bench and a launch list.
mkdir -p gpurun_out
echo "== unit tests (gemm, attention)"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "gemm or attention_core" --no-header -p no:cacheprovider > gpurun_out/t_unit.log 2>&1; rc=$?; echo "rc=$rc"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/t_unit.log | tail -20
if [ -n "$EXPERIMENT" ]; then echo "== experiment: $EXPERIMENT"; env $EXPERIMENT timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "gemm" --no-header -p no:cacheprovider 2>&1 | grep -E "^(FAILED|ERROR)|passed|failed" | tail -12; fi
if [ $rc -ne 0 ] && [ -z "$CONTINUE" ]; then exit 0; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu -k "not gemm and not attention_core" --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/t_gpu.log | tail -12
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
