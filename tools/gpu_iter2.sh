#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_gpu.log
echo "== bench (PDL on)"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench.json; grep -o '"roofline.*"cpu_baseline' gpurun_out/bench.json | cut -c1-900; grep -o '"e2e".*' gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench (PDL off)"; FLEXDM_PDL=0 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench_nopdl.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
