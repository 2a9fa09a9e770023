#!/bin/bash
# quick GPU check: parity suite + bench (tf32 / bf16 lidar, tf32 fusion; no cpu baseline)
set -u
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
$T python bench.py --steps 200 --warmup 20 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "bench exit $?"
$T python bench.py --steps 200 --warmup 20 --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err
$T python bench.py --steps 200 --warmup 20 --workload fusion --no-cpu-baseline > gpurun_out/bench_fusion_fp16.json 2> gpurun_out/bench_fusion_fp16.err
$T python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
python - <<'PY'
import json
for f in ("bench_tf32","bench_bf16","bench_fp16","bench_fusion_fp16"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"], "e2e", round(d["e2e"]["value"]), "roof", d["roofline"]["frac"], "hbm", d["hbm_roofline"]["frac"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
