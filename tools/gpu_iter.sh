#!/bin/bash
# Iteration visit: parity tests, short bench, launch list, optional ncu --set full (NCU_K regex, NCU_S skip, NCU_C count).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
if [ -n "$NCU_K" ]; then
  echo "== ncu full $NCU_K"; tools/gpu_ncu_full.sh "$NCU_K" ${NCU_S:-0} ${NCU_C:-8} full_iter
  ncu -i gpurun_out/full_iter.ncu-rep --page details --csv > gpurun_out/full_iter.details.csv 2>/dev/null
  ncu -i gpurun_out/full_iter.ncu-rep --page raw --csv > gpurun_out/full_iter.raw.csv 2>/dev/null
fi
du -sm gpurun_out
