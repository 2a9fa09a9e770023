#!/bin/bash
# product build check: whole GPU suite (with per-test timeout), then bench lidar / fusion / tf32
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 120 2>&1 | tail -n 3
for a in "" "--workload fusion" "--precision tf32" "--workload fusion_layer"; do
timeout 120 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 $a 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$a', round(d['value']), round(d['ms_per_step']*1e3,2), round(d['one_batch_in_flight']['ms_per_step']*1e3,2), {k: round(v*1e3,1) for k,v in d['stage_ms'].items()}, round(d['roofline']['frac'],3))"
done
