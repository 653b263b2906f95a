#!/bin/bash
# Evidence for the last build of round 2 (gate-bits build): smoke, both bench arms, ncu launch list with DRAM bytes, ncu --set full of the
# FFN launches whose epilogues changed (FFN 1 with the gate words, {dW2, dHid} reading them) plus the heaviest forward GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== ncu launch list + dram bytes"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
# 37 GEMM launches per step; step 3 (the measured one) starts at GEMM index 111: encoder 2, then per block QKV, O, FFN1, FFN2 -> FFN 1 of block 0 is 111 + 4
echo "== ncu full gemm (measured step, first 8 forward GEMMs + the first FFN backward group)"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tf32" -s 111 -c 24 -o gpurun_out/full_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_gemm.log 2>&1; echo "rc=$?"
ncu -i gpurun_out/full_gemm.ncu-rep --page raw --csv > gpurun_out/full_gemm.raw.csv 2>/dev/null
rm -f gpurun_out/full_gemm.ncu-rep
du -sm gpurun_out
