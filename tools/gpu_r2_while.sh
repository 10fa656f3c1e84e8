#!/bin/bash
# round 2: the Jacobi iteration as one WHILE graph (device-side convergence): tests + A/B on one bond and on the sweep
out=gpurun_out/${1:-r3i}; mkdir -p $out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_mps.py tests/test_gpu_fullsize.py tests/test_gpu_chain.py -q -m gpu -x > $out/pytest.log 2>&1
tail -n 3 $out/pytest.log | cut -c1-300
QB200_DEBUG=1 timeout 300 python tools/prof_bond.py 1024 > $out/bond.log 2> $out/bond.err; tail -1 $out/bond.log; grep -c sweep $out/bond.err; grep sweep $out/bond.err | tail -3
for o in 1 0; do QB200_SVD_WHILE=$o python bench.py --steps 3 --warmup 2 --no-sliced --no-expect --no-cpu-baseline > $out/bench_w$o.log 2> $out/bench_w$o.err; python -c "
import json
d=json.loads(open('$out/bench_w$o.log').read().strip().splitlines()[-1]); print($o, round(d['value'],4), d['detail']['step_wall_ms'], round(d['e2e']['value'],4), d['roofline']['jacobi_sweeps_per_svd'], d.get('parity_check'))"; done
