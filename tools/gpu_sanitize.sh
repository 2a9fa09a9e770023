#!/bin/bash
# compute-sanitizer passes over the GPU suite: memcheck (whole suite), then racecheck / synccheck / initcheck on the kernels' tests
set -u
mkdir -p gpurun_out
for tool in racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fusion_layer.py tests/test_gpu_train.py tests/test_gpu_pipeline.py tests/test_gpu_determinism.py -q -m gpu -p no:cacheprovider -x > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/sanitize_$tool.log | tail -3
done
