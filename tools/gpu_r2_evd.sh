#!/bin/bash
# round 2: evd kernel with W in registers (cross steps) + short rotation chain: A/B on one bulk bond, GPU tests
out=gpurun_out/${1:-r2v}; mkdir -p $out
timeout 300 python tools/ab_bond.py 1024 > $out/ab_wreg.log 2>&1
QB200_EVD_WREG=0 timeout 300 python tools/ab_bond.py 1024 > $out/ab_nowreg.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_fullsize.py tests/test_gpu_chain.py -q -m gpu -x > $out/pytest.log 2>&1
for f in ab_wreg ab_nowreg; do echo "== $f"; tail -n 1 $out/$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k in ('rep0', 'rep1'):
    r = d[k]; print(r['kept'], r['sweeps'], r['dw'], r['lam_head'], {k: v for k, v in r['phases_ms'].items()})"; done
tail -n 5 $out/pytest.log | cut -c1-300
