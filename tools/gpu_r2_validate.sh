#!/bin/bash
# round-2 validation on a B200: full GPU test-suite, smoke, both bench arms (short)
out=gpurun_out/${1:-r2v}
mkdir -p $out
(time timeout 1500 python -m pytest tests -m gpu -x -q --durations=12) > $out/pytest.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
timeout 500 python bench.py --steps 5 --warmup 3 > $out/bench.log 2> $out/bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $out/bench_ref.log 2> $out/bench_ref.err
tail -n 12 $out/pytest.log; tail -n 2 $out/smoke.log; tail -c 400 $out/bench.err; head -c 600 $out/bench.log
