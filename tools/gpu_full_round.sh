#!/bin/bash
# One GPU visit: parity tests, bench line, reference arm, ncu launch list, ncu --set full of the GEMM and attention kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --durations=8 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_bench.log
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "rc=$?"; cat gpurun_out/bench_ref.json
if [ -z "$NO_NCU_FULL" ]; then
echo "== ncu full gemm"; tools/gpu_ncu_full.sh "gemm_tf32" 150 50 full_gemm
echo "== ncu full attention"; tools/gpu_ncu_full.sh "attention_" 24 8 full_attn
fi
