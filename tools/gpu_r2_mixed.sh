#!/bin/bash
# round 2: mixed-precision Jacobi (FP32 stage A + FP64 stage B): correctness and per-phase timing on one bulk bond
out=gpurun_out/r2q; mkdir -p $out
QB200_SVD_MIXED=1 QB200_DEBUG=1 timeout 300 python tools/ab_bond.py 1024 > $out/ab_mixed.log 2> $out/ab_mixed.err
timeout 300 python tools/ab_bond.py 1024 > $out/ab_base.log 2>&1
QB200_SVD_MIXED=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_fullsize.py -q -m gpu -x > $out/pytest_mixed.log 2>&1
grep "stage A\|sweep" $out/ab_mixed.err | head -40
for f in ab_mixed ab_base; do echo "== $f"; tail -n 1 $out/$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k in ('rep0', 'rep1'):
    r = d[k]; print(r['kept'], r['sweeps'], r['dw'], r['lam_head'], {k: v for k, v in r['phases_ms'].items()})"; done
tail -n 5 $out/pytest_mixed.log | cut -c1-300
