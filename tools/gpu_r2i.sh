#!/bin/bash
# round 2, pass i: HostPipeline tests + the bench's e2e legs
set -u
mkdir -p gpurun_out
T="timeout 400"
$T python -m pytest tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -x > gpurun_out/pytest_pipe.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/pytest_pipe.log
$T python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_pipe.json").read().strip().splitlines()[-1])
    print("tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"])
    e=d["e2e"]; print("e2e", round(e["value"]), e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], "fp32pts", round(e["fp32_points_input"]["value"]), "full d2h", round(e["full_result_d2h"]["value"]), e["host_numa"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_pipe.err").read()[-3000:])
PY
$T python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --workload fusion_layer > gpurun_out/bench_pipe_fl.json 2> gpurun_out/bench_pipe_fl.err; echo "bench fl exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_pipe_fl.json").read().strip().splitlines()[-1])
    print("FL tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4))
    e=d["e2e"]; print("FL e2e", round(e["value"]), e["h2d_bytes_per_step"], e["d2h_bytes_per_step"], "fp32pts", round(e["fp32_points_input"]["value"]), "full d2h", round(e["full_result_d2h"]["value"]))
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_pipe_fl.err").read()[-3000:])
PY
