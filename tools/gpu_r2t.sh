#!/bin/bash
# kCrossBatch 256 (product) vs 512 (variant), same box, alternating
set -u
for i in 1 2; do
for v in default cb512; do
  if [ $v = default ]; then unset P3P_LIB; else export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_$v.so; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), round(d['ms_per_step']*1e3,2), round(d['one_batch_in_flight']['ms_per_step']*1e3,2), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})"
done; done
unset P3P_LIB
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -n 2
