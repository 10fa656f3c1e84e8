#!/bin/bash
# round 2: A/B of the second-generation Jacobi update / cross-Gram kernels (operand sums formed once per chunk)
out=gpurun_out/r2p; mkdir -p $out
timeout 120 python tools/dmma_patterns.py > $out/dmma_patterns.log 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py -q -m gpu -x > $out/pytest_a.log 2>&1
QB200_UPDATE_V2=0 QB200_GRAM_V2=0 timeout 300 python tools/ab_bond.py 1024 > $out/ab_old.log 2>&1
QB200_GRAM_V2=0 timeout 300 python tools/ab_bond.py 1024 > $out/ab_upd2.log 2>&1
QB200_UPDATE_V2=0 timeout 300 python tools/ab_bond.py 1024 > $out/ab_gram2.log 2>&1
timeout 300 python tools/ab_bond.py 1024 > $out/ab_both.log 2>&1
cat $out/dmma_patterns.log
tail -n 5 $out/pytest_a.log | cut -c1-300
for f in ab_old ab_upd2 ab_gram2 ab_both; do echo "== $f"; tail -n 1 $out/$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['rep1']; print(r['kept'], r['sweeps'], r['lam_head'], {k: v for k, v in r['phases_ms'].items()})"; done
