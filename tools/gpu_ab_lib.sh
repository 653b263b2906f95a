#!/bin/bash
# A/B of two BUILDS of the engine on one box, interleaved twice: the in-tree libflexdm_mfp.so against csrc/build/libflexdm_mfp_base.so
# (a copy of the previous build; build/ is git-ignored but travels with gpurun).  GPU suite on the new build first.
mkdir -p gpurun_out
echo "== pytest -m gpu (new build)"; timeout 900 python -m pytest tests -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t_gpu.log; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | cut -c1-250 | head -20
NEW=flex_dm_b200/libflexdm_mfp.so; BASE=flex_dm_b200/csrc/build/libflexdm_mfp_base.so
cp $NEW /tmp/new.so
for v in new base new base; do
  if [ $v = base ]; then cp $BASE $NEW; else cp /tmp/new.so $NEW; fi
  timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-tfrecord --no-other-configs --no-e2e "$@" > gpurun_out/bench_ab_lib.json 2> gpurun_out/bench_ab.err || tail -5 gpurun_out/bench_ab.err
  python - "$v" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/bench_ab_lib.json").read().strip().splitlines()[-1])
    print("%-8s ms/step %.4f" % (sys.argv[1], d["ms_per_step"]))
except Exception as e:
    print(sys.argv[1], "bench parse failed:", e)
PY
done
cp /tmp/new.so $NEW
