#!/bin/bash
# the driver's multi-GPU commands at N = 2: own arm and reference arm under torchrun
set -u
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err ) 2>&1 | grep real; echo "2-GPU bench exit $?"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err ) 2>&1 | grep real; echo "2-GPU reference exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read().strip().splitlines()[-1])
print("tiles/s", round(d["value"]), "us/step", round(d["ms_per_step"]*1e3,2), "e2e", round(d["e2e"]["value"]), "full d2h", round(d["e2e"]["full_result_d2h"]["value"]), d["e2e"].get("host_numa"), d["clocks"])
for s in d["sub_results"]: print("   ", {k: (round(v,4) if isinstance(v,float) else v) for k,v in s.items() if k not in ("what","conv_roofline")})
r=json.loads(open("gpurun_out/bench_2gpu_ref.json").read().strip().splitlines()[-1])
print({k: r[k] for k in ("impl","value","unit","n_gpus","cpu_baseline")})
PY
tail -3 gpurun_out/bench_2gpu.err
