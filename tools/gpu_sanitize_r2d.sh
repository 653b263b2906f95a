#!/bin/bash
# compute-sanitizer over the kernels that changed after the round-2b sanitizer pass (mask_corrupt rewrite, loss kernel, attention backward
# with O rows by TMA, encoder GEMM with the fused LayerNorm), then the whole GPU suite once more (second full run of the final build).
mkdir -p gpurun_out
K='mask_corrupt or mask_for_test or forward_loss_backward or attention_core_backward or loss_layer_standalone or packed'
echo "== memcheck"; timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_packed_columns.py -q -m gpu --no-header -p no:cacheprovider -k "$K" > gpurun_out/r2d_memcheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r2d_memcheck.log | tail -3
echo "== racecheck"; timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu --no-header -p no:cacheprovider -k "mask_corrupt or mask_for_test or (attention_core_backward and tcgen05) or loss_layer_standalone" > gpurun_out/r2d_racecheck.log 2>&1; echo "rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/r2d_racecheck.log | tail -3
echo "== pytest -m gpu (full, again)"; timeout 900 python -m pytest tests -x -q -m gpu --no-header -p no:cacheprovider > gpurun_out/t_gpu_again.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_gpu_again.log
