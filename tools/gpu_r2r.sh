#!/bin/bash
# voxelizer phase timeline per chunk index (timeline build in a variant library: the product library is untouched)
set -u
mkdir -p gpurun_out pixelspointspolygons_b200/variants
python tools/build_variant.py tl -DP3P_TIMELINE > gpurun_out/build_tl.log 2>&1 || tail gpurun_out/build_tl.log
export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so
timeout 300 python tools/timeline.py 16 100000 > gpurun_out/vox_tl_chunks.txt 2>&1; echo "exit $?"
cat gpurun_out/vox_tl_chunks.txt
