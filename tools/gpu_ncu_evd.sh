#!/bin/bash
out=gpurun_out/${1:-r2w}; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:jacobi_evd_kernel -s 200 -c 2 -o $out/r2_evd_wreg -f python tools/prof_bond.py 1024 > $out/ncu_evd.log 2>&1
tail -3 $out/ncu_evd.log
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_fullsize.py tests/test_gpu_chain.py -q -m gpu -x > $out/pytest.log 2>&1
tail -n 3 $out/pytest.log | cut -c1-300
