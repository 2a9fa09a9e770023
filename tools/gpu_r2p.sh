#!/bin/bash
# voxelizer shape variants (P3P_LIB selects the build): parity subset + bench
set -u
mkdir -p gpurun_out
for v in default v512 v384; do
  if [ $v = default ]; then unset P3P_LIB; else export P3P_LIB=$PWD/pixelspointspolygons_b200/variants/libp3p_$v.so; fi
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider 2>&1 | tail -n 2
  timeout 300 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 > gpurun_out/bench_vox_$v.json 2>/dev/null
  timeout 300 python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 --points 400000 > gpurun_out/bench_vox400_$v.json 2>/dev/null
  python - <<PY
import json
for f in ("bench_vox_$v","bench_vox400_$v"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, "tiles/s", round(d["value"]), "us/step", round(d["ms_per_step"]*1e3,2), "one", round(d["one_batch_in_flight"]["ms_per_step"]*1e3,2), {k: round(x*1e3,1) for k,x in d["stage_ms"].items() if x})
PY
done
