#!/bin/bash
# ncu evidence for round 1 (second session): full captures of the changed hot kernels + launch lists.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
NCU="ncu --clock-control none"
# 1. launch list of one bulk bond (all kernels, serialised)
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r1b_bond_launches.csv python tools/prof_bond.py 1024 > gpurun_out/ncu1.log 2>&1
# 2. full captures: update (3M), cross Gram (3M), evd (17 warps)
timeout 300 $NCU --set full --import-source on -k regex:jacobi_update_kernel -s 200 -c 2 -o gpurun_out/r1b_update -f python tools/prof_bond.py 1024 > gpurun_out/ncu2.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:jacobi_gram_kernel -s 200 -c 2 -o gpurun_out/r1b_gram -f python tools/prof_bond.py 1024 > gpurun_out/ncu3.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:jacobi_evd_kernel -s 200 -c 2 -o gpurun_out/r1b_evd -f python tools/prof_bond.py 1024 > gpurun_out/ncu4.log 2>&1
# 3. ComplexF32 GEMM and the FP64 GEMM (3M) at 4096^3
timeout 300 $NCU --set full --import-source on -k regex:gemm_c64_kernel -s 4 -c 1 -o gpurun_out/r1b_gemm_c64 -f python tools/time_gemm_c64.py > gpurun_out/ncu5.log 2>&1
timeout 300 $NCU --set full --import-source on -k regex:gemm_c128_kernel -s 8 -c 1 -o gpurun_out/r1b_gemm_c128 -f python tools/time_gemm.py > gpurun_out/ncu6.log 2>&1
# 4. HBM helpers at 512 MiB: dram bytes + duration for every launch of the gather / scale / norm kernels
timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:'gather|scale_mode|sumsq' --csv --log-file gpurun_out/r1b_hbm_launches.csv python tools/time_hbm.py 4096 > gpurun_out/ncu7.log 2>&1
# 5. launch list of the bench command itself: 30000 launches from inside the timed sweep
timeout 900 $NCU --metrics gpu__time_duration.sum -s 160000 -c 30000 --csv --log-file gpurun_out/r1b_bench_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-sliced > gpurun_out/ncu8.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv; tail -3 gpurun_out/ncu8.log
