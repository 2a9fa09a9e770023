#!/bin/bash
# round 2: fusion-layer convolution (TMA + tcgen05) first light: parity, bench
set -u
mkdir -p gpurun_out
T="timeout 300"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc"; tail -n 3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
timeout 200 python -m pytest tests/test_gpu_fusion_layer.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_fl.log 2>&1; echo "pytest fusion_layer exit $?"; tail -n 25 gpurun_out/pytest_fl.log
$T python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
$T python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-sub-results --min-seconds 1 --workload fusion_layer > gpurun_out/bench_fl.json 2> gpurun_out/bench_fl.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_fl.json").read().strip().splitlines()[-1])
    print("fusion_layer tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"], "e2e", round(d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_fl.err").read()[-3000:])
PY
$T python - <<'PY'
import json, torch, sys
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda:0")
w = bench.Workload(dev, 0, "fusion_layer", "fp16", 16, 100000, 64, 2)
print("conv roofline", json.dumps(bench.conv_roofline(w, json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {})))
PY
$T ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 2 -c 1 -o gpurun_out/prof_conv -f \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline --no-sub-results --min-seconds 0.01 --workload fusion_layer > gpurun_out/ncu_conv.log 2>&1; echo "ncu exit $?"
