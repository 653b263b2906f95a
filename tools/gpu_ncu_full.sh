#!/bin/bash
# ncu --set full of selected kernels of the bench step.  usage: gpu_ncu_full.sh <kernel-regex> <skip> <count> <outname>
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$1" -s "$2" -c "$3" -o "gpurun_out/$4" -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > "gpurun_out/$4.log" 2>&1
echo "rc=$?"; tail -3 "gpurun_out/$4.log"; ls -la gpurun_out/*.ncu-rep
