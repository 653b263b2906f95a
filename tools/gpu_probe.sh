#!/bin/bash
mkdir -p gpurun_out
echo "== store microbench"; for cfg in "2048 4 0" "2048 16 0" "2048 4 128" "2048 16 128" "64 4 128" "64 16 0"; do timeout 60 tools/microbench/store_bw $cfg 2>&1 | tail -1; done | tee gpurun_out/store_bw.txt
echo "== ncu attention"; timeout 600 ncu --set full --clock-control none --import-source on -k "regex:attention_" -s 24 -c 5 -o gpurun_out/attn7 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/attn7.log 2>&1; echo rc=$?
ncu -i gpurun_out/attn7.ncu-rep --page raw --csv > gpurun_out/attn7.raw.csv 2>/dev/null
ncu -i gpurun_out/attn7.ncu-rep --page source --csv --kernel-id :::1 > gpurun_out/attn7_fwd.source.csv 2>/dev/null
ncu -i gpurun_out/attn7.ncu-rep --page source --csv --kernel-id :::5 > gpurun_out/attn7_bwd.source.csv 2>/dev/null
echo "== ncu gemm qkv"; timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gemm_tf32" -s 170 -c 1 -o gpurun_out/qkv7 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/qkv7.log 2>&1; echo rc=$?
ncu -i gpurun_out/qkv7.ncu-rep --page raw --csv > gpurun_out/qkv7.raw.csv 2>/dev/null
ncu -i gpurun_out/qkv7.ncu-rep --page source --csv > gpurun_out/qkv7.source.csv 2>/dev/null
rm -f gpurun_out/qkv7.ncu-rep
ls -la gpurun_out; du -sm gpurun_out
