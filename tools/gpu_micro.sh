#!/bin/bash
mkdir -p gpurun_out
echo "== microbench"; for cfg in "32 16 4" "32 48 3" "32 48 4" "64 32 6" "2048 48 4"; do timeout 60 tools/microbench/l2_fill $cfg 2>&1 | tail -n +3 | sed -n '1p;4p'; done | tee gpurun_out/l2_fill.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
