#!/bin/bash
# round-2 profiles: ncu launch lists of one bench command per workload + one full capture per kernel (1 GPU; numbers printed
# under ncu are never bench values), then the bench lines themselves (CUDA events, not under ncu)
set -u
mkdir -p gpurun_out
T="timeout 300"
Q="--no-graph --no-cpu-baseline --no-sub-results --min-seconds 0.01 --in-flight 1"
$T ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/launches_fusion_layer.csv \
   python bench.py --steps 5 --warmup 2 $Q --workload fusion_layer > gpurun_out/ncu_launch_fl.log 2>&1; echo "launch list (fusion_layer) exit $?"
$T ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 5 --warmup 2 $Q > gpurun_out/ncu_launch.log 2>&1; echo "launch list (lidar) exit $?"
for k in pfn_tc voxelize_kernel; do
  $T ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/prof_$k -f \
     python bench.py --steps 3 --warmup 1 $Q > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
for k in conv3x3_tc patch_embed_tc; do
  $T ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k -f \
     python bench.py --steps 3 --warmup 1 $Q --workload fusion_layer > gpurun_out/ncu_$k.log 2>&1; echo "ncu $k exit $?"
done
$T ncu --set full --clock-control none -k regex:pfn_tc -s 6 -c 1 -o gpurun_out/prof_pfn_tc_tf32 -f \
   python bench.py --steps 3 --warmup 1 $Q --precision tf32 > gpurun_out/ncu_pfn_tf32.log 2>&1; echo "ncu pfn tf32 exit $?"
$T ncu --set full --clock-control none -k regex:las_ -s 2 -c 2 -o gpurun_out/prof_las -f \
   python tools/pipe_probe.py > gpurun_out/ncu_las.log 2>&1; echo "ncu las exit $?"
ls -la gpurun_out/*.ncu-rep
for wl in lidar fusion fusion_layer; do
  $T python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --workload $wl > gpurun_out/bench_r02_$wl.json 2> gpurun_out/bench_r02_$wl.err; echo "bench $wl exit $?"
done
$T python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --precision tf32 > gpurun_out/bench_r02_lidar_tf32.json 2>/dev/null; echo "bench tf32 exit $?"
