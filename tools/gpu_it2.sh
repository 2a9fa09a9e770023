#!/bin/bash
# quick iteration: parity suite, fp16 + tf32 bench, then timelines of both kernels
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/pytest_gpu.log
for prec in fp16 tf32; do
timeout 300 python bench.py --steps 200 --warmup 20 --precision $prec --no-cpu-baseline > gpurun_out/bench_$prec.json 2> gpurun_out/bench_$prec.err; echo "bench exit $?"
done
python - <<'PY'
import json
for f in ("bench_fp16","bench_tf32"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"], "e2e", round(d["e2e"]["value"]), "roof", round(d["roofline"]["frac"],3), d["clocks"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
if [ "${TL:-1}" = "1" ]; then
P3P_EXTRA_NVCC_FLAGS=-DP3P_TIMELINE python -m pixelspointspolygons_b200.build --force > gpurun_out/build_tl.log 2>&1 || tail gpurun_out/build_tl.log
timeout 300 python tools/timeline.py 16 100000 > gpurun_out/vox_tl.txt 2>&1; echo "exit $?"
timeout 300 python tools/pfn_timeline.py fp16 > gpurun_out/pfn_tl_fp16.txt 2>&1; echo "exit $?"
cat gpurun_out/vox_tl.txt
tail -n 14 gpurun_out/pfn_tl_fp16.txt
fi
