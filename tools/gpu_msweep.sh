#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
: > gpurun_out/msweep.jsonl
for m in 4 16 32 128 256 512; do
  timeout 300 python bench.py --steps 20 --warmup 3 --max-points-per-voxel $m --no-cpu-baseline 2>/dev/null | tail -n 1 >> gpurun_out/msweep.jsonl
done
timeout 300 python bench.py --steps 20 --warmup 3 --precision fp32 --no-cpu-baseline 2>/dev/null | tail -n 1 >> gpurun_out/msweep.jsonl
python - <<'PY'
import json
for l in open("gpurun_out/msweep.jsonl"):
    try: d=json.loads(l)
    except Exception: print("ERR", l[:200]); continue
    c=d["config"]
    print("M", c["max_points_per_voxel"], c["precision"], "| us/step", round(d["ms_per_step"]*1e3,1), "tiles/s", round(d["value"]), "kernel", d["roofline"]["kernel"], "stage", {k: round(v*1e3,1) for k,v in d["stage_ms"].items() if v}, "kept", d["roofline"]["kept_points"])
PY
