#!/bin/bash
out=gpurun_out/r2j; mkdir -p $out
timeout 600 python tools/gpu_debug_r2h.py > $out/canon.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_chain.py tests/test_gpu_c64.py tests/test_gpu_semantics.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -q -m gpu > $out/pytest_a.log 2>&1
timeout 900 python -m pytest tests/test_gpu_configs.py -q -m gpu -k "config3" > $out/pytest_c3.log 2>&1
timeout 400 python bench.py --steps 3 --warmup 2 --no-sliced --no-expect --no-cpu-baseline > $out/bench.log 2> $out/bench.err
echo "== canon"; tail -n 12 $out/canon.log | cut -c1-300; tail -n 15 $out/pytest_a.log | cut -c1-300; tail -n 15 $out/pytest_c3.log | cut -c1-300; tail -c 300 $out/bench.err; head -c 300 $out/bench.log
