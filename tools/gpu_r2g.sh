#!/bin/bash
set -u
mkdir -p gpurun_out
T="timeout 300"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc"; tail -n 2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED -- stopping"; tail -n 30 gpurun_out/smoke.log; exit 1; fi
$T python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"], "roof", d["roofline"]["frac"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
}
B="--steps 50 --warmup 10 --no-cpu-baseline --no-sub-results --min-seconds 1"
$T python bench.py $B > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; summ bench_fp16
$T python bench.py $B --precision tf32 > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; summ bench_tf32
$T python bench.py $B --points 400000 > gpurun_out/bench_400k.json 2> gpurun_out/bench_400k.err; summ bench_400k
$T python bench.py $B --points 200000 > gpurun_out/bench_200k.json 2> gpurun_out/bench_200k.err; summ bench_200k
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so $T python tools/pfn_timeline.py fp16 > gpurun_out/pfn_tl_fp16.txt 2>&1; echo "tl exit $?"; tail -n 16 gpurun_out/pfn_tl_fp16.txt
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so $T python tools/timeline.py 16 400000 > gpurun_out/vox_tl_400k.txt 2>&1; echo "vox tl exit $?"; tail -n 14 gpurun_out/vox_tl_400k.txt
