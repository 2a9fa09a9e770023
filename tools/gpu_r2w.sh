#!/bin/bash
# a PFN variant library against the product build: probe, bench (alternating), parity
set -u
V=${1:-unroll}
export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_$V.so
timeout 60 python tools/epi4_probe.py | tail -2; echo "probe exit $?"
for i in 1 2; do
for v in default $V; do
  if [ $v = default ]; then unset P3P_LIB; else export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_$v.so; fi
  timeout 120 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), round(d['ms_per_step']*1e3,2), round(d['one_batch_in_flight']['ms_per_step']*1e3,2), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()})"
done; done
