#!/bin/bash
# Development loop: GEMM / attention unit tests first (a protocol bug must not hang the whole suite), then the GPU suite, then a short bench.
# usage: gpu_quick.sh [bench args...]
mkdir -p gpurun_out
echo "== gemm unit tests"; timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "gemm or attention_core" --no-header -p no:cacheprovider -x > gpurun_out/t_unit.log 2>&1; rc=$?; echo "rc=$rc"; tail -3 gpurun_out/t_unit.log
if [ $rc -ne 0 ]; then grep -E "^E |Error|timeout|trap" gpurun_out/t_unit.log | head -20; exit 0; fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu.log; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | cut -c1-250 | head -20
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-tfrecord --no-other-configs "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("ms/step %.4f  e2e %.4f  gemm %.4f ms (%d launches, frac %.3f)  attn %.4f ms  launches/step %.1f  loss0 %s" % (
        d["ms_per_step"], d["e2e"]["ms_per_step"], r["gemm_ms_per_step"], r["gemm_launches_per_step"], r["frac"], r["attention"]["ms_per_step"],
        d["gpu_launches"] / d["steps"], d["config"]["loss_step0"]))
except Exception as e:
    print("bench parse failed:", e)
    print(open("gpurun_out/bench_quick.err").read()[-1500:])
PY
