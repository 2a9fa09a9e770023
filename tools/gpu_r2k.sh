#!/bin/bash
set -u
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "patch_embed or fusion or pipeline" > gpurun_out/pytest_pe.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_pe.log
for wl in fusion fusion_layer; do
$T python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 --workload $wl 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', 'ms/step', d['ms_per_step'], 'one', d['one_batch_in_flight']['ms_per_step'], 'e2e', d['e2e']['value'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"patch_embed" -c 4 python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline --no-sub-results --min-seconds 0.01 --workload fusion 2>&1 | grep -E "patch_embed|gpu__time" | head -8
