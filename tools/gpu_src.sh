#!/bin/bash
# source-level ncu of the N-th launch matching a kernel regex: usage gpu_src.sh <regex> <skip> <name>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$1" -s $2 -c 1 -o gpurun_out/$3 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/$3.log 2>&1; echo rc=$?
ncu -i gpurun_out/$3.ncu-rep --page raw --csv > gpurun_out/$3.raw.csv 2>/dev/null
ncu -i gpurun_out/$3.ncu-rep --page source --csv > gpurun_out/$3.source.csv 2>/dev/null
rm -f gpurun_out/$3.ncu-rep
