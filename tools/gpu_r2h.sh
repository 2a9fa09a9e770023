#!/bin/bash
set -u
mkdir -p gpurun_out
T="timeout 400"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc"; tail -n 2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED -- stopping"; tail -n 30 gpurun_out/smoke.log; exit 1; fi
$T python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
$T python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
    print("tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "burst", d["burst"]["ms_per_step"], "stage_ms", d["stage_ms"], "roof", d["roofline"]["frac"], d["clocks"])
    e=d["e2e"]; print("e2e", round(e["value"]), e["h2d_bytes_per_step"], "fp32pts", round(e["fp32_points_input"]["value"]), "full d2h", round(e["full_result_d2h"]["value"]), e["host_numa"])
    print("cpu", d["cpu_baseline"])
    for s in d["sub_results"]: print("  sub", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k != "conv_roofline"}, s.get("conv_roofline", {}).get("frac"))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_default.err").read()[-3000:])
PY
$T python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -c 600 gpurun_out/bench_ref.json
