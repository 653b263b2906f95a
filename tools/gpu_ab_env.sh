#!/bin/bash
# A/B of an environment switch on the current build, interleaved twice: usage gpu_ab_env.sh VAR=val [bench args...]
mkdir -p gpurun_out
sw=$1; shift
for v in "A=1" "$sw" "A=1" "$sw"; do
  env $v timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-tfrecord --no-other-configs --no-e2e "$@" > gpurun_out/bench_ab_env.json 2> gpurun_out/bench_ab.err || tail -5 gpurun_out/bench_ab.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_ab_env.json").read().strip().splitlines()[-1])
    print("%-24s ms/step %.4f" % (sys.argv[1], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "bench parse failed:", e)
PY
done
