#!/bin/bash
# voxelizer change check: parity + fuzz, bench (twice), per-chunk timeline from a variant build
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -n 2
for i in 1 2; do
for n in 100000 400000; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 --points $n 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['points_per_tile'], round(d['value']), round(d['ms_per_step']*1e3,2), round(d['one_batch_in_flight']['ms_per_step']*1e3,2), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})"
done; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 --workload fusion 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fusion', round(d['value']), round(d['ms_per_step']*1e3,2))"
python tools/build_variant.py tl -DP3P_TIMELINE > gpurun_out/build_tl.log 2>&1 || tail gpurun_out/build_tl.log
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so timeout 300 python tools/timeline.py 16 100000 > gpurun_out/vox_tl_chunks.txt 2>&1; echo "exit $?"
head -13 gpurun_out/vox_tl_chunks.txt; tail -26 gpurun_out/vox_tl_chunks.txt | awk '{print $1,$2,$3,$11,$12,$13,$15}'
