#!/bin/bash
# round 2: persistent streaming thin kernel: parity tests, per-node timing of one slice, sliced timing A/B
out=gpurun_out/${1:-r3e}; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tn.py -q -m gpu -x > $out/pytest.log 2>&1
tail -n 4 $out/pytest.log | cut -c1-300
QB200_DEBUG_TN=1 python tools/probe_sliced_nodes.py 40 6 24 > $out/nodes.log 2> $out/nodes.err
grep -A400 "rep 1" $out/nodes.err | sort -k22 -n -r | head -10
python tools/time_sliced.py 40 6 24 > $out/sliced_stream.log 2>&1; tail -1 $out/sliced_stream.log
QB200_GEMM_STREAM=0 python tools/time_sliced.py 40 6 24 > $out/sliced_nostream.log 2>&1; tail -1 $out/sliced_nostream.log
python tools/time_sliced.py 40 7 24 64 > $out/sliced_d7.log 2>&1; tail -1 $out/sliced_d7.log
