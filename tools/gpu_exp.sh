#!/bin/bash
set -u
mkdir -p gpurun_out
P3P_EXTRA_NVCC_FLAGS="-DP3P_TIMELINE -DP3P_EXP_NOFENCE" python -m pixelspointspolygons_b200.build --force > gpurun_out/build_tl.log 2>&1 || tail gpurun_out/build_tl.log
timeout 300 python tools/pfn_timeline.py fp16 > gpurun_out/pfn_tl_fp16_nofence.txt 2>&1; echo "exit $?"
tail -n 14 gpurun_out/pfn_tl_fp16_nofence.txt
python -m pixelspointspolygons_b200.build --force > gpurun_out/build.log 2>&1
bash tools/gpu_prof.sh pfn_tc prof_pfn
