#!/bin/bash
# density sweep of BASELINE config 5 (lidar_density_ablation shape): points per tile at M = 64, M at N = 100k; B = 16, 1 GPU
set -u
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
for n in 10000 25000 50000 100000 200000 400000; do
  timeout 300 python bench.py --steps 100 --warmup 10 --points $n --no-cpu-baseline 2>/dev/null | tail -n 1 >> gpurun_out/sweep.jsonl
done
for m in 16 32 128 256; do
  timeout 300 python bench.py --steps 100 --warmup 10 --max-points-per-voxel $m --no-cpu-baseline 2>/dev/null | tail -n 1 >> gpurun_out/sweep.jsonl
done
timeout 300 python bench.py --steps 100 --warmup 10 --batch 8 --workload fusion --no-cpu-baseline 2>/dev/null | tail -n 1 >> gpurun_out/sweep.jsonl
timeout 300 python bench.py --steps 100 --warmup 10 --batch 32 --no-cpu-baseline 2>/dev/null | tail -n 1 >> gpurun_out/sweep.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    try:
        d=json.loads(l)
    except Exception as e:
        print("ERR", l[:200]); continue
    c=d["config"]
    print(c["workload"][:28], "B", c["tiles_per_gpu"], "N", c["points_per_tile"], "M", c["max_points_per_voxel"], "| us/step", round(d["ms_per_step"]*1e3,1), "tiles/s", round(d["value"]), "Mpts/s", round(d["mpoints_per_s"]), "hbm", round(d["hbm_roofline"]["frac"],3), "kernel", d["roofline"]["kernel"], "stage", {k: round(v*1e3,1) for k,v in d["stage_ms"].items() if v})
PY
