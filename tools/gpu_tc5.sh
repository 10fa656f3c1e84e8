#!/bin/bash
# ComplexF32 GEMM on tcgen05 (opt-in): parity suite + timing + accuracy vs K
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 QB200_C64_TCGEN05=1 QB200_C64_TCGEN05_CHECK=1
timeout 70 python -m pytest tests/test_gpu_c64.py -m gpu -q > gpurun_out/tc5_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/tc5_pytest.log
unset QB200_C64_TCGEN05_CHECK
timeout 60 python tools/time_gemm_c64.py > gpurun_out/tc5_gemm.log 2>&1
timeout 60 python tools/tc5_accuracy.py > gpurun_out/tc5_acc_tc5.log 2>&1
tail -12 gpurun_out/tc5_pytest.log | cut -c1-200; cat gpurun_out/tc5_gemm.log gpurun_out/tc5_acc_tc5.log
