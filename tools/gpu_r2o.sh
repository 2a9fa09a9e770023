#!/bin/bash
# round 2: compute-sanitizer over the training tests and smoke(), then the density / M sweeps (1 GPU)
set -u
mkdir -p gpurun_out
T="timeout 900"
$T compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_train.py -q -m gpu -x -p no:cacheprovider -k "sparse or m16 or second_step" > gpurun_out/memcheck_train.log 2>&1; echo "memcheck train exit $?"; tail -n 6 gpurun_out/memcheck_train.log
$T compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py smoke > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck smoke exit $?"; tail -n 4 gpurun_out/memcheck_smoke.log
: > gpurun_out/sweep.jsonl
for n in 10000 25000 50000 100000 200000 400000; do
  timeout 300 python bench.py --steps 20 --warmup 5 --points $n --no-cpu-baseline --no-sub-results --min-seconds 1 2>/dev/null | tail -n 1 >> gpurun_out/sweep.jsonl
done
for m in 4 16 32 128 256 512; do
  timeout 300 python bench.py --steps 20 --warmup 5 --max-points-per-voxel $m --no-cpu-baseline --no-sub-results --min-seconds 1 2>/dev/null | tail -n 1 >> gpurun_out/sweep.jsonl
done
timeout 300 python bench.py --steps 20 --warmup 5 --precision fp32 --no-cpu-baseline --no-sub-results --min-seconds 1 2>/dev/null | tail -n 1 >> gpurun_out/sweep.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    try:
        d=json.loads(l)
    except Exception as e:
        print("ERR", l[:200]); continue
    c=d["config"]
    print("N", c["points_per_tile"], "M", c["max_points_per_voxel"], c["precision"], "| us/step", round(d["ms_per_step"]*1e3,1), "one", round(d["one_batch_in_flight"]["ms_per_step"]*1e3,1), "tiles/s", round(d["value"]), "Mpts/s", round(d["mpoints_per_s"]), "hbm", round(d["hbm_roofline"]["frac"],3), d["roofline"]["kernel"], {k: round(v*1e3,1) for k,v in d["stage_ms"].items() if v})
PY
