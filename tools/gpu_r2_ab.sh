#!/bin/bash
out=gpurun_out/r2o; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_fullsize.py tests/test_gpu_semantics.py tests/test_gpu_chain.py tests/test_gpu_tn.py -q -m gpu > $out/pytest_a.log 2>&1
QB200_QR_REFINE=0 timeout 300 python tools/ab_bond.py 1024 > $out/ab_norefine.log 2>&1
timeout 300 python tools/ab_bond.py 1024 > $out/ab_refine.log 2>&1
QB200_QR_REFINE=0 timeout 400 python bench.py --steps 4 --warmup 2 --no-sliced --no-expect --no-cpu-baseline > $out/bench_norefine.log 2> $out/bench_norefine.err
timeout 600 python bench.py --steps 4 --warmup 2 > $out/bench.log 2> $out/bench.err
tail -n 12 $out/pytest_a.log | cut -c1-300
for f in ab_norefine ab_refine; do echo "== $f"; tail -n 1 $out/$f.log | cut -c1-700; done
for f in bench_norefine bench; do echo "== $f"; head -c 330 $out/$f.log; echo; done; tail -c 300 $out/bench.err
