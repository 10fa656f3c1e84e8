#!/bin/bash
# Round-1 (second session) GPU check: 3M complex product A/B, parity tests, bench.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -c "import sys; sys.path.insert(0, '.'); import qrochet_b200 as qb; c = qb.Context(0); print('dmma', c.dmma_peak_tflops(), 'hmma tf32/bf16', c.hmma_peak_tflops())" > gpurun_out/peaks.log 2>&1
( for u in 0 1; do QB200_UPDATE_3M=$u QB200_GEMM_3M=$u timeout 120 python tools/ab_bond.py 1024; done ) > gpurun_out/ab_bond.log 2>&1
( for u in 0 1; do echo "GEMM_3M=$u"; QB200_GEMM_3M=$u timeout 120 python tools/time_gemm.py; done ) > gpurun_out/ab_gemm.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1
cat gpurun_out/peaks.log; tail -3 gpurun_out/ab_bond.log; tail -25 gpurun_out/ab_gemm.log; tail -25 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/bench.log
