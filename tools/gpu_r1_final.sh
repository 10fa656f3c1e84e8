#!/bin/bash
# Final validation of the session: full GPU suite, smoke(), configs 1-3 wall times, the bench line.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final.log 2>&1
echo "bench exit $?" >> gpurun_out/bench_final.log
timeout 300 python tools/run_configs.py > gpurun_out/configs_final.log 2>&1
tail -4 gpurun_out/pytest_gpu_final.log; tail -2 gpurun_out/smoke_final.log; tail -3 gpurun_out/configs_final.log | cut -c1-1500; tail -2 gpurun_out/bench_final.log | cut -c1-6000
