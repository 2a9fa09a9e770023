#!/bin/bash
# default bench (1 GPU), reference arm, and both under torchrun on all visible GPUs
set -u
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
if [ "$N" -gt 1 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; echo "ref N=$N exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 200 --warmup 20 --workload fusion --batch 8 --no-cpu-baseline > gpurun_out/bench_fusion_n$N.json 2> gpurun_out/bench_fusion_n$N.err; echo "fusion N=$N exit $?"
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_default.json")+glob.glob("gpurun_out/bench_ref*.json")+glob.glob("gpurun_out/bench_n*.json")+glob.glob("gpurun_out/bench_fusion_n*.json")):
    try:
        lines=[l for l in open(f).read().strip().splitlines() if l.startswith("{")]
        d=json.loads(lines[-1])
        print(f, "lines", len(lines), "n_gpus", d.get("n_gpus"), "value", round(d["value"],1), d["unit"], "ms/step", d.get("ms_per_step"), "e2e", round(d["e2e"]["value"],1), "impl", d.get("impl"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "ERR", e)
PY
