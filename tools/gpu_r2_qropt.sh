#!/bin/bash
# round 2: optimistic QR pass (no per-panel host round trip): tests + A/B on one bond and on the sweep
out=gpurun_out/${1:-r3h}; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_fullsize.py tests/test_gpu_chain.py tests/test_gpu_configs.py -q -m gpu -x > $out/pytest.log 2>&1
tail -n 3 $out/pytest.log | cut -c1-300
timeout 300 python tools/ab_bond.py 1024 > $out/ab_opt.log 2>&1
QB200_QR_OPTIMISTIC=0 timeout 300 python tools/ab_bond.py 1024 > $out/ab_noopt.log 2>&1
for f in ab_opt ab_noopt; do echo "== $f"; tail -n 1 $out/$f.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
r = d['rep1']; print(r['kept'], r['sweeps'], r['dw'], r['lam_head'], {k: v for k, v in r['phases_ms'].items()})"; done
for o in 1 0; do QB200_QR_OPTIMISTIC=$o python bench.py --steps 3 --warmup 2 --no-sliced --no-expect --no-cpu-baseline > $out/bench_o$o.log 2> $out/bench_o$o.err; python -c "
import json
d=json.loads(open('$out/bench_o$o.log').read().strip().splitlines()[-1]); print($o, round(d['value'],4), d['detail']['step_wall_ms'], round(d['e2e']['value'],4))"; done
