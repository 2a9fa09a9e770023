#!/bin/bash
# in-flight sweep + the full default bench line (sub-results incl. the training step)
set -u
mkdir -p gpurun_out
T="timeout 900"
for f in 1 2 3 4; do
  $T python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 --in-flight $f > gpurun_out/bench_if$f.json 2>/dev/null; echo "in-flight $f exit $?"
done
for f in 2 3; do
  $T python bench.py --steps 20 --warmup 5 --no-sub-results --no-cpu-baseline --min-seconds 1 --in-flight $f --workload fusion > gpurun_out/bench_fusion_if$f.json 2>/dev/null; echo "fusion in-flight $f exit $?"
done
( time $T python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real; echo "default bench exit $?"
python - <<'PY'
import json
for f in ("bench_if1","bench_if2","bench_if3","bench_if4","bench_fusion_if2","bench_fusion_if3","bench_default"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "tiles/s", round(d["value"]), "us/step", round(d["ms_per_step"]*1e3,2), "one", round(d["one_batch_in_flight"]["ms_per_step"]*1e3,2), "e2e", round(d["e2e"]["value"]), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        for s in d.get("sub_results", []):
            print("    ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.items() if k not in ("what", "conv_roofline")})
    except Exception as e:
        print(f, "ERR", e)
PY
