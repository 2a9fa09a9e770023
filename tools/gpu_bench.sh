#!/bin/bash
# bench + ncu passes on the GPU box (one GPU).
set -u
mkdir -p gpurun_out
T="timeout 600"
$T python bench.py --steps 200 --warmup 20 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
$T python bench.py --steps 200 --warmup 20 --no-graph --no-cpu-baseline > gpurun_out/bench_tf32_nograph.json 2> gpurun_out/bench_tf32_nograph.err
$T python bench.py --steps 200 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
$T ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 5 --warmup 2 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:pfn_tc -s 10 -c 2 -o gpurun_out/prof_pfn \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_pfn.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:rank_kernel -s 20 -c 2 -o gpurun_out/prof_rank \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_rank.log 2>&1
tail -c 3000 gpurun_out/bench_tf32.json; tail -n 5 gpurun_out/bench_tf32.err
tail -c 1500 gpurun_out/bench_tf32_nograph.json; tail -c 1500 gpurun_out/bench_bf16.json
ls -la gpurun_out
