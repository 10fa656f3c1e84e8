#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python tools/time_gemm_c64.py > gpurun_out/gemm_c64.log 2>&1
timeout 120 python tools/ab_bond.py 1024 > gpurun_out/ab_bond2.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sliced > gpurun_out/bench_f.log 2>&1
tail -8 gpurun_out/pytest_gpu.log; cat gpurun_out/gemm_c64.log; cut -c1-700 gpurun_out/ab_bond2.log
python -c "
import json
l=[x for x in open('gpurun_out/bench_f.log') if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['phases_ms_one_bulk_bond'])
"
