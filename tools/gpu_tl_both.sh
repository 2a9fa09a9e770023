#!/bin/bash
# parity suite, then phase timelines of both kernels (library rebuilt with -DP3P_TIMELINE on the box)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
P3P_EXTRA_NVCC_FLAGS=-DP3P_TIMELINE python -m pixelspointspolygons_b200.build --force > gpurun_out/build_tl.log 2>&1 || tail gpurun_out/build_tl.log
timeout 300 python tools/timeline.py 16 100000 > gpurun_out/vox_tl.txt 2>&1; echo "exit $?"
timeout 300 python tools/pfn_timeline.py fp16 > gpurun_out/pfn_tl_fp16.txt 2>&1; echo "exit $?"
cat gpurun_out/vox_tl.txt
tail -n 25 gpurun_out/pfn_tl_fp16.txt
