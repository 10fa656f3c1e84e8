#!/bin/bash
# round 2: mixed-precision Jacobi (FP32 stage A on tcgen05 + FP64 stage B): correctness and per-phase timing on one bulk bond
out=gpurun_out/${1:-r2t}; mkdir -p $out
QB200_SVD_MIXED=1 QB200_LP_UPDATE=check QB200_DEBUG=1 QB200_C64_TCGEN05_CHECK=1 timeout 300 python tools/ab_bond.py 1024 > $out/ab_check.log 2> $out/ab_check.err
QB200_SVD_MIXED=1 QB200_DEBUG=1 timeout 300 python tools/ab_bond.py 1024 > $out/ab_mixed.log 2> $out/ab_mixed.err
QB200_SVD_MIXED=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_fullsize.py -q -m gpu -x > $out/pytest_mixed.log 2>&1
grep "check\|stage A\|sweep" $out/ab_check.err | head -15
for f in ab_check ab_mixed; do echo "== $f"; tail -n 1 $out/$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k in ('rep0', 'rep1'):
    r = d[k]; print(r['kept'], r['sweeps'], r['dw'], r['lam_head'], {k: v for k, v in r['phases_ms'].items()})"; done
tail -n 5 $out/pytest_mixed.log | cut -c1-300
if [ -n "$2" ]; then
QB200_SVD_MIXED=1 python bench.py --steps 3 --warmup 1 > $out/bench_mixed.log 2> $out/bench_mixed.err; tail -c 600 $out/bench_mixed.err; python - <<EOP
import json
d = json.loads(open("$out/bench_mixed.log").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["detail"]["step_wall_ms"], d["e2e"]["value"], d.get("parity_check"))
print(d["roofline"]["frac"], d["roofline"].get("jacobi_sweeps_per_svd"))
for k, v in d["roofline"]["kernels_one_sweep_all_streams"].items(): print(k, v)
EOP
fi
