#!/bin/bash
# ncu-only GPU visit: launch list of one step + --set full of a few GEMM / attention launches, exported to CSV on the box
# (gpurun_out/ is capped at 64 MiB: keep reports small).
mkdir -p gpurun_out
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_bench.log
echo "== ncu full gemm"; tools/gpu_ncu_full.sh "gemm_tf32" ${GEMM_SKIP:-168} ${GEMM_COUNT:-14} full_gemm
echo "== ncu full attention"; tools/gpu_ncu_full.sh "attention_" 24 4 full_attn
for r in full_gemm full_attn; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/$r.details.csv 2>/dev/null
done
ls -la gpurun_out; du -sm gpurun_out
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/full_gemm.ncu-rep; fi
