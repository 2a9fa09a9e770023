#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke(), reference arm, own arm (1 GPU)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err ) 2>&1 | grep real; echo "reference arm exit $?"
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real; echo "bench exit $?"
python - <<'PY'
import json
r=json.loads(open("gpurun_out/bench_final_ref.json").read().strip().splitlines()[-1])
d=json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("reference", round(r["value"],2), r["unit"], r["cpu_baseline"]["cores"], "cores |", "ours", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ratio e2e", round(d["e2e"]["value"]/r["value"]), "roofline", round(d["roofline"]["frac"],3), "hbm", round(d["hbm_roofline"]["frac"],3), "launches", d["gpu_launches"], d["clocks"], d["cpu_baseline"])
PY
# per-chunk voxelizer timeline (variant build on the box)
mkdir -p pixelspointspolygons_b200/variants
python tools/build_variant.py tl -DP3P_TIMELINE > gpurun_out/build_tl.log 2>&1 || tail gpurun_out/build_tl.log
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so timeout 120 python tools/timeline.py 16 100000 > gpurun_out/vox_tl_chunks.txt 2>&1; echo "timeline exit $?"
