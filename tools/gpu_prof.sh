#!/bin/bash
# ncu full capture of one kernel of the bench step: usage gpu_prof.sh <kernel-regex> <out-name> [bench args...]
set -u
mkdir -p gpurun_out
K=$1; O=$2; shift 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -o gpurun_out/$O -f \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline "$@" > gpurun_out/ncu_$O.log 2>&1
echo "ncu exit $?"; tail -n 3 gpurun_out/ncu_$O.log | cut -c1-300
