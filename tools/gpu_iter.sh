#!/bin/bash
mkdir -p gpurun_out
echo "== gemm unit tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "gemm" --no-header -p no:cacheprovider -x > gpurun_out/t_unit.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/t_unit.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|timeout|trap" gpurun_out/t_unit.log | head -20; exit 0; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu.log
for v in "A=1" "$1"; do echo "== bench [$v]"; env $v timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_ab.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_$v.json | head -1; grep -o '"gemm_ms_per_step[^,]*' gpurun_out/bench_$v.json; grep -o '"e2e".*' gpurun_out/bench_$v.json | cut -c1-60; tail -2 gpurun_out/bench_ab.err; done
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
