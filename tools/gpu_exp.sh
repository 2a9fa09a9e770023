#!/bin/bash
set -u
mkdir -p gpurun_out
for exp in "-DP3P_REG_MMA=40 -DP3P_REG_FRONT=96 -DP3P_REG_EPI=80" "-DP3P_REG_MMA=32 -DP3P_REG_FRONT=104 -DP3P_REG_EPI=80"; do
P3P_EXTRA_NVCC_FLAGS="$exp" python -m pixelspointspolygons_b200.build --force > gpurun_out/build_exp.log 2>&1 || tail gpurun_out/build_exp.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "pfn or encode or fusion" 2>&1 | tail -1
for prec in fp16 tf32; do
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --precision $prec > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err
python - "$exp $prec" <<'PY'
import json,sys
d=json.loads(open("gpurun_out/bench_exp.json").read().strip().splitlines()[-1])
print("EXP[%s]"%sys.argv[1], "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"])
PY
done
done
