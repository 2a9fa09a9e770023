#!/bin/bash
# one optimisation iteration on the GPU box: parity suite, bench (default fp16, tf32, fusion), ncu launch list + full captures
set -u
mkdir -p gpurun_out
T="timeout 600"
$T python -m pytest tests -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 6 gpurun_out/pytest_gpu.log
$T python bench.py --steps 200 --warmup 20 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; echo "bench exit $?"
$T python bench.py --steps 200 --warmup 20 --precision tf32 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err
$T python bench.py --steps 200 --warmup 20 --workload fusion --no-cpu-baseline > gpurun_out/bench_fusion_fp16.json 2> gpurun_out/bench_fusion_fp16.err
$T ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 5 --warmup 2 --no-graph --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:pfn_tc -s 6 -c 1 -o gpurun_out/prof_pfn -f \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_pfn.log 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:voxelize_kernel -s 6 -c 1 -o gpurun_out/prof_vox -f \
   python bench.py --steps 3 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_vox.log 2>&1
python - <<'PY'
import json
for f in ("bench_fp16","bench_tf32","bench_fusion_fp16"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "tiles/s", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "stage_ms", d["stage_ms"], "e2e", round(d["e2e"]["value"]), "roof", d["roofline"]["frac"], "hbm", d["hbm_roofline"]["frac"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-2000:])
PY
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
