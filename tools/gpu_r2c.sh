#!/bin/bash
# round 2: parity + new bench line + voxelizer timeline
set -u
mkdir -p gpurun_out
T="timeout 200"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc"; tail -n 6 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
$T python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
$T python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
    print("tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "burst", d["burst"]["ms_per_step"], "stage_ms", d["stage_ms"], "e2e", round(d["e2e"]["value"]), round(d["e2e"]["full_result_d2h"]["value"]), "roof", d["roofline"]["frac"], d["clocks"], d["cpu_baseline"])
    for s in d["sub_results"]: print("  sub", s)
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_default.err").read()[-3000:])
PY
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so $T python tools/timeline.py > gpurun_out/vox_tl.txt 2>&1; echo "vox tl exit $?"; tail -n 14 gpurun_out/vox_tl.txt
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so $T python tools/timeline.py 16 400000 > gpurun_out/vox_tl_400k.txt 2>&1; echo "vox tl exit $?"; tail -n 14 gpurun_out/vox_tl_400k.txt
