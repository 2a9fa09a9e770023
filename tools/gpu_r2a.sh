#!/bin/bash
# round 2, pass a: PFN with two accumulator stages per tile + counting voxelizer -- parity, bench, timeline, variants
set -u
mkdir -p gpurun_out
T="timeout 150"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke exit $rc"; tail -n 6 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED -- stopping"; exit 1; fi
$T python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_gpu.log
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"], "e2e", round(d["e2e"]["value"]), "roof", d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
}
$T python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; summ bench_fp16
$T python bench.py --steps 200 --warmup 20 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; summ bench_tf32
for v in "$@"; do
  P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_$v.so $T python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; summ bench_$v
done
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so $T python tools/pfn_timeline.py fp16 > gpurun_out/pfn_tl_fp16.txt 2>&1; echo "tl exit $?"; tail -n 16 gpurun_out/pfn_tl_fp16.txt
P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_tl.so $T python tools/timeline.py > gpurun_out/vox_tl.txt 2>&1; echo "vox tl exit $?"; tail -n 20 gpurun_out/vox_tl.txt
$T ncu --set full --clock-control none --import-source on -k regex:pfn_tc -s 6 -c 1 -o gpurun_out/prof_pfn -f \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_pfn.log 2>&1; echo "ncu exit $?"
$T ncu --set full --clock-control none --import-source on -k regex:voxelize_kernel -s 6 -c 1 -o gpurun_out/prof_vox -f \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1; echo "ncu exit $?"
