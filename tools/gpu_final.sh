#!/bin/bash
# Round-end evidence on one GPU: parity tests, smoke, bench (both arms), ncu launch list with DRAM bytes, ncu --set full of the
# GEMM / attention kernels (raw + details pages exported to CSV on the box; the reports themselves are dropped).
# usage: gpu_final.sh [tag]      (artefacts land in gpurun_out/; copy the summaries to profiles/<tag>_*)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --durations=5 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -10 gpurun_out/t_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-300
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "== bench 3xTF32"; timeout 600 python bench.py --gemm-impl 2 --steps 30 --warmup 5 --no-cpu-baseline --no-tfrecord --no-other-configs > gpurun_out/bench_3xtf32.json 2> gpurun_out/bench_3xtf32.err; echo "rc=$?"
echo "== ncu launch list + dram bytes"; timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
# the step has 37 GEMM launches (value leg: step 0 + 2 warm-up steps before the measured one): skip 3 x 37 and take one step's forward GEMMs
echo "== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:gemm_tf32" -s 111 -c 14 -o gpurun_out/full_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_gemm.log 2>&1; echo "rc=$?"
echo "== ncu full attention"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:attention_" -s 27 -c 2 -o gpurun_out/full_attn -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/full_attn.log 2>&1; echo "rc=$?"
for r in full_gemm full_attn; do
  ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
  ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/$r.details.csv 2>/dev/null
done
rm -f gpurun_out/full_gemm.ncu-rep gpurun_out/full_attn.ncu-rep
python tools/microbench/copy_bw_by_size.py > gpurun_out/copy_bw.log 2>&1
du -sm gpurun_out
