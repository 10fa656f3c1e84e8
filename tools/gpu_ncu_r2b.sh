#!/bin/bash
# round 2, second batch of ncu captures: thin GEMM kernel (one slice of the sliced benchmark), tcgen05 stage-A update
out=gpurun_out/${1:-r3d}; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_mixed_svd.py -q -m gpu -x > $out/pytest_mixed.log 2>&1; tail -3 $out/pytest_mixed.log | cut -c1-300
QB200_TN_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_c128_thin_kernel -s 60 -c 6 -o $out/r2_gemm_thin -f python tools/probe_sliced_nodes.py 40 6 24 > $out/ncu_thin.log 2>&1; tail -2 $out/ncu_thin.log
QB200_SVD_MIXED=1 QB200_SVD_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:lp_update_tc5_kernel -s 100 -c 2 -o $out/r2_lp_update_tc5 -f python tools/prof_bond.py 1024 > $out/ncu_lp.log 2>&1; tail -2 $out/ncu_lp.log
