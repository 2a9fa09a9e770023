#!/bin/bash
# four-group epilogue (variant library) against the product build: bench on the same box, alternating; then parity
set -u
mkdir -p gpurun_out
for i in 1 2; do
for v in default epi4; do
  if [ $v = default ]; then unset P3P_LIB; else export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_$v.so; fi
  for wl in lidar fusion; do
  timeout 120 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 --workload $wl 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v $wl', round(d['value']), round(d['ms_per_step']*1e3,2), round(d['one_batch_in_flight']['ms_per_step']*1e3,2), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()}, round(d['roofline']['frac'],3))"
  done
done; done
export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_epi4.so
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_fusion_layer.py tests/test_gpu_pipeline.py -q -m gpu -x -p no:cacheprovider --timeout 60 2>&1 | tail -n 6
