#!/bin/bash
# Mixed-precision Gram A/B + parity tests + sweep bench.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( for u in 0 1; do QB200_GRAM_LOWP=$u QB200_DEBUG=1 timeout 120 python tools/ab_bond.py 1024; done ) > gpurun_out/ab_lowp.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for u in 0 1; do
  QB200_GRAM_LOWP=$u timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sliced > gpurun_out/bench_lowp$u.log 2>&1
done
grep -h "sweep\|^{" gpurun_out/ab_lowp.log | cut -c1-1500 | head -60; tail -15 gpurun_out/pytest_gpu.log
for u in 0 1; do python -c "
import json,sys
l=[x for x in open('gpurun_out/bench_lowp$u.log') if x.startswith('{')]
d=json.loads(l[-1]); print('LOWP=$u', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['phases_ms_one_bulk_bond'])
"; done
