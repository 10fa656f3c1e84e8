#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python tools/time_gemm_c64.py > gpurun_out/gemm_c64.log 2>&1
timeout 120 python tools/time_hbm.py 4096 > gpurun_out/hbm4096.log 2>&1
tail -12 gpurun_out/pytest_gpu.log; cat gpurun_out/gemm_c64.log gpurun_out/hbm4096.log
