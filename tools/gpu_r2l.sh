#!/bin/bash
# round 2, training step: new tests first, then the whole GPU suite, then the timing of the training step
set -u
mkdir -p gpurun_out
T="timeout 900"
$T python -m pytest tests/test_gpu_train.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_train.log 2>&1; echo "pytest train exit $?"; tail -n 40 gpurun_out/pytest_train.log
$T python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu.log
$T python tools/time_train.py > gpurun_out/time_train.json 2> gpurun_out/time_train.err; echo "time_train exit $?"; cat gpurun_out/time_train.json; tail -5 gpurun_out/time_train.err
$T ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/time_train.py 16 100000 fused-only > gpurun_out/ncu_train.log 2>&1; echo "ncu train exit $?"
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/launches_train.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
d=collections.defaultdict(list)
for r in rows[1:]: d[r[ki][:70]].append(float(r[vi].replace(',','')))
for k,v in d.items(): print('  %-70s n=%3d mean_us=%9.1f'%(k,len(v),sum(v)/len(v)/1e3))
PY
