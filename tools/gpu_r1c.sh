#!/bin/bash
# ComplexF32 path tests + timing, and sweep A/B of the 3M complex product.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_c64.py tests/test_gpu_kernels.py tests/test_gpu_chain.py -m gpu -q --durations=5 > gpurun_out/pytest_c64.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_c64.log
timeout 120 python tools/time_gemm_c64.py > gpurun_out/gemm_c64.log 2>&1
for u in 0 1; do
  QB200_UPDATE_3M=$u QB200_GEMM_3M=$u timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sliced > gpurun_out/bench_3m$u.log 2>&1
done
tail -40 gpurun_out/pytest_c64.log; cat gpurun_out/gemm_c64.log
for u in 0 1; do python -c "
import json,sys
l=[x for x in open('gpurun_out/bench_3m$u.log') if x.startswith('{')]
d=json.loads(l[-1]); print('3M=$u', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'])
"; done
