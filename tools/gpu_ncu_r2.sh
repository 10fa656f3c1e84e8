#!/bin/bash
# ncu evidence for round 2: launch list of one bulk bond and of the bench command, full captures of the Jacobi kernels,
# the GEMMs and the sliced executor.  (A number printed by a run under ncu is never a bench value.)
out=gpurun_out/r2ncu; mkdir -p $out
export PYTHONUNBUFFERED=1 QB200_SVD_GRAPH=0 QB200_TN_GRAPH=0   # plain launches: -s / -c skip counts stay meaningful
NCU="ncu --clock-control none"
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $out/r2_bond_launches.csv python tools/prof_bond.py 1024 > $out/ncu1.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:jacobi_update_kernel -s 200 -c 2 -o $out/r2_update -f python tools/prof_bond.py 1024 > $out/ncu2.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:jacobi_gram_kernel -s 200 -c 2 -o $out/r2_gram -f python tools/prof_bond.py 1024 > $out/ncu3.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:jacobi_evd_kernel -s 200 -c 2 -o $out/r2_evd -f python tools/prof_bond.py 1024 > $out/ncu4.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:gemm_c128_kernel -s 8 -c 1 -o $out/r2_gemm_c128 -f python tools/time_gemm.py > $out/ncu5.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:gemm_c64_tc5_kernel -s 4 -c 1 -o $out/r2_gemm_c64_tc5 -f python tools/time_gemm_c64.py > $out/ncu6.log 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum -s 100000 -c 30000 --csv --log-file $out/r2_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sliced --no-expect > $out/ncu7.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file $out/r2_sliced_launches.csv python tools/probe_sliced.py 40 6 24 > $out/ncu8.log 2>&1
ls -la $out/*.ncu-rep $out/*.csv; tail -n 3 $out/ncu7.log | cut -c1-300
