#!/bin/bash
# round 2: full parity suite with the new tests, conv with three issuers, fusion_layer bench
set -u
mkdir -p gpurun_out
T="timeout 300"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc"; tail -n 2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
$T python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
$T python - <<'PY'
import json, torch, sys, os
sys.path.insert(0, ".")
import bench
dev = torch.device("cuda:0")
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
for B in (16, 8, 32):
    w = bench.Workload(dev, 0, "fusion_layer", "fp16", B, 100000, 64, 2)
    print("B", B, "conv roofline", json.dumps(bench.conv_roofline(w, peaks)))
    w.time_steps(20); steps, ms = w.time_for(0.5, chunk=64)
    print("B", B, "fusion_layer step ms", ms / steps, "tiles/s", B * steps / ms * 1e3)
    del w; torch.cuda.empty_cache()
PY
