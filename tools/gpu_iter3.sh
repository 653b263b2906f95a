#!/bin/bash
mkdir -p gpurun_out
echo "== tensor store microbench"; for cfg in "768 32 32 2 4" "768 32 32 4 4" "768 32 128 2 1" "768 64 32 2 4" "768 256 4 2 4" "768 256 16 2 4" "256 32 32 2 4" "256 256 16 2 4"; do timeout 60 tools/microbench/tstore_bw $cfg 2>&1 | tail -1; done | tee gpurun_out/tstore_bw.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-330 gpurun_out/bench.json; grep -o '"gemm_ms_per_step[^,]*' gpurun_out/bench.json; grep -o '"attention": {"ms_per_step[^,]*' gpurun_out/bench.json;  grep -o '"e2e".*' gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
