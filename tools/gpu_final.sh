#!/bin/bash
# Round-end evidence: parity tests, smoke, bench (with CPU baseline), reference arm, launch list with DRAM bytes, ncu --set full of the
# dominant kernels (raw/details CSV exported on the box; reports kept small).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --durations=5 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "== ncu launch list + dram bytes"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_bench.log
echo "== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tf32" -s 168 -c 12 -o gpurun_out/full_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_gemm.log 2>&1; echo "rc=$?"
echo "== ncu full attention"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attention_" -s 27 -c 2 -o gpurun_out/full_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_attn.log 2>&1; echo "rc=$?"
for r in full_gemm full_attn; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/$r.details.csv 2>/dev/null
done
rm -f gpurun_out/full_gemm.ncu-rep
du -sm gpurun_out
