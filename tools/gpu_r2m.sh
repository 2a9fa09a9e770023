#!/bin/bash
# ncu captures of the training kernels (B = 16 x 100 k)
set -u
mkdir -p gpurun_out
T="timeout 600"
for k in train_stats1 train_back2 train_forward; do
$T ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/prof_$k -f python tools/time_train.py 16 100000 fused-only > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
