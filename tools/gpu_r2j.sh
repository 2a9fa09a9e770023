#!/bin/bash
# round 2, pass j: lanes (batches in flight) -- tests + the default bench line
set -u
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_pipe.log 2>&1; echo "pytest exit $?"; tail -n 5 gpurun_out/pytest_pipe.log
$T python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
    print("tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "burst", d["burst"]["ms_per_step"], "one-in-flight", d["one_batch_in_flight"]["ms_per_step"], "stage_ms", d["stage_ms"], "roof", d["roofline"]["frac"], d["clocks"])
    e=d["e2e"]; print("e2e", round(e["value"]), e["h2d_bytes_per_step"], "fp32pts", round(e["fp32_points_input"]["value"]), "full d2h", round(e["full_result_d2h"]["value"]), e["host_numa"])
    print("cpu", d["cpu_baseline"])
    for s in d["sub_results"]: print("  sub", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k != "conv_roofline"}, s.get("conv_roofline", {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_default.err").read()[-3000:])
PY
