#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-sliced > gpurun_out/bench_g.log 2>&1
python -c "
import json
l=[x for x in open('gpurun_out/bench_g.log') if x.startswith('{')]
d=json.loads(l[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['step_wall_ms'], d['config']['jacobi_sweeps_per_svd'], d['gpu_launches'])
"
