#!/bin/bash
out=gpurun_out/r2n; mkdir -p $out
timeout 900 python tools/gpu_debug_r2h.py > $out/canon.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_chain.py tests/test_gpu_semantics.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py tests/test_gpu_c64.py -q -m gpu > $out/pytest_a.log 2>&1
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -k "config3" > $out/pytest_c3.log 2>&1
timeout 300 python tools/run_configs.py > $out/configs.log 2>&1
echo "== canon"; tail -n 12 $out/canon.log | cut -c1-300; tail -n 25 $out/pytest_a.log | cut -c1-300; tail -n 8 $out/pytest_c3.log | cut -c1-300; tail -n 5 $out/configs.log | cut -c1-600
