#!/bin/bash
# A/B bench of an environment switch: usage gpu_ab.sh VAR=1
mkdir -p gpurun_out
echo "== unit tests with $1"; env $1 timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -k "gemm or golden or forward_loss" > gpurun_out/t_ab.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_ab.log
for v in "" "$1"; do echo "== bench [$v]"; env $v timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_ab.json | head -1; grep -o '"gemm_ms_per_step[^,]*' gpurun_out/bench_ab.json; tail -2 gpurun_out/bench_ab.err; done
echo "== ncu launch list [$1]"; env $1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_ab.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ab.log 2>&1; echo "rc=$?"
