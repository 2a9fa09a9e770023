#!/bin/bash
# First-light sequence on the GPU box: each stage under its own timeout so that a hung kernel cannot block the rest.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; local t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$t" "$@" > "gpurun_out/$name.log" 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
}
run vox      600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "voxelizer" -p no:cacheprovider
run fp32     600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "features_and_canvas and fp32" -p no:cacheprovider
run tf32     600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "features_and_canvas and tf32" -p no:cacheprovider
run bf16     600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "features_and_canvas and bf16" -p no:cacheprovider
run rest     900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "not voxelizer and not features_and_canvas" -p no:cacheprovider
run smoke    300 python __graft_entry__.py smoke
