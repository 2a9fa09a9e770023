#!/bin/bash
set -u
export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_epi4.so
timeout 60 python tools/epi4_probe.py | tail -3; echo "probe exit $?"
timeout 120 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('epi4/64 lidar', round(d['value']), round(d['ms_per_step']*1e3,2), round(d['one_batch_in_flight']['ms_per_step']*1e3,2), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})"
