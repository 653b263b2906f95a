#!/bin/bash
# Short round check: parity tests, smoke, default bench (with the TFRecord input-pipeline legs), optionally the reference arm and a
# compute-sanitizer pass over the kernels added last (usage: gpu_round.sh [ref] [sanitize]).
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 600 python -m pytest tests -q -rxXw -m gpu --no-header -p no:cacheprovider --durations=5 > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gpu.log; grep -E "^FAILED|^XPASS|^XFAIL|^E  |UserWarning: TF32 operand rounding" gpurun_out/t_gpu.log | cut -c1-300 | head -30
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for arg in "$@"; do
  if [ "$arg" = "ref" ]; then echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "rc=$?"; cat gpurun_out/bench_ref.json; fi
  if [ "$arg" = "sanitize" ]; then
    echo "== compute-sanitizer memcheck (context token, document gather)"
    timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_context.py tests/test_gpu_input_pipeline.py tests/test_gpu_train_flow.py -q -m gpu --no-header -p no:cacheprovider -k "gradients or device_cached or train_py or ctx_id_shuffled" > gpurun_out/sanitizer.log 2>&1; echo "rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer.log | tail -4
  fi
done
nproc
