#!/bin/bash
out=gpurun_out/r2g; mkdir -p $out
timeout 120 python tools/gpu_debug_r2f.py tn > $out/tn_graph.log 2>&1
QB200_TN_GRAPH=0 timeout 120 python tools/gpu_debug_r2f.py tn > $out/tn_nograph.log 2>&1
timeout 300 python tools/gpu_debug_r2f.py qr > $out/qr.log 2>&1
QB200_NO_CHOLQR=1 timeout 300 python tools/gpu_debug_r2f.py qr > $out/qr_nochol.log 2>&1
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_semantics.py tests/test_gpu_tn.py tests/test_gpu_fullsize.py -q -m gpu > $out/pytest_rest.log 2>&1
timeout 500 python bench.py --steps 3 --warmup 2 > $out/bench.log 2> $out/bench.err
QB200_SVD_GRAPH=0 timeout 500 python bench.py --steps 3 --warmup 2 --no-sliced --no-expect --no-cpu-baseline > $out/bench_nograph.log 2> $out/bench_nograph.err
for f in tn_graph tn_nograph qr qr_nochol; do echo "== $f"; tail -n 12 $out/$f.log | cut -c1-250; done; tail -n 15 $out/pytest_rest.log | cut -c1-250; tail -c 300 $out/bench.err
