#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider 2>&1 | tail -2
timeout 300 python bench.py --steps 200 --warmup 20 --workload fusion --no-cpu-baseline > gpurun_out/bench_fusion_fp16.json 2> gpurun_out/bench_fusion_fp16.err; echo "exit $?"; tail -n 3 gpurun_out/bench_fusion_fp16.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_fusion_fp16.json").read().strip().splitlines()[-1])
print("fusion tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"], "e2e", round(d["e2e"]["value"]), "hbm", round(d["hbm_roofline"]["frac"],4))
PY
