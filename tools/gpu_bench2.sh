#!/bin/bash
# default bench (1 GPU), reference arm, and the same under torchrun on all visible GPUs
set -u
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -c 1800 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; tail -c 600 gpurun_out/bench_ref.json
if [ "$N" -gt 1 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N exit $?"; tail -c 1500 gpurun_out/bench_n$N.json; tail -n 5 gpurun_out/bench_n$N.err
fi
