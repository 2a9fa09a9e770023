#!/bin/bash
set -u
mkdir -p gpurun_out
P3P_EXTRA_NVCC_FLAGS=-DP3P_TIMELINE python -m pixelspointspolygons_b200.build --force > gpurun_out/build_tl.log 2>&1 || tail gpurun_out/build_tl.log
timeout 300 python tools/pfn_timeline.py tf32 > gpurun_out/pfn_tl_tf32.txt 2>&1; echo "exit $?"
timeout 300 python tools/pfn_timeline.py bf16 > gpurun_out/pfn_tl_bf16.txt 2>&1; echo "exit $?"
tail -n 30 gpurun_out/pfn_tl_tf32.txt
