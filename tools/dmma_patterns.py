"""DMMA issue-pattern micro-benchmarks (libqrochet_b200_diag.so): what a complex-multiply DMMA stream can reach, and
what the operand sums of the 3M product cost next to it."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import qrochet_b200 as qb
from qrochet_b200 import _capi as capi
ctx = qb.Context(0)
d = capi.diag()
out = (C.c_double * 9)()
capi.check(ctx.h, d.qb200_bench_dmma_patterns(ctx.h, out))
names = ["independent acc, fixed operands", "complex pattern, register operands", "complex pattern, A from LDS.128"]
print("TFLOP/s at 8 / 16 / 32 warps per SM")
for m in range(3):
    print(f"{names[m]:40s}", " ".join(f"{out[m * 3 + o]:7.2f}" for o in range(3)))
out6 = (C.c_double * 6)()
capi.check(ctx.h, d.qb200_bench_dmma_3m(ctx.h, out6))
names = ["3M: sum by DADD per fragment", "3M: sum from a shared plane (LDS.64)", "3M: no third operand (DMMA + LDS.128)"]
print("executed DMMA TFLOP/s at 8 / 16 warps per SM")
for m in range(3):
    print(f"{names[m]:40s}", " ".join(f"{out6[m * 2 + o]:7.2f}" for o in range(2)))
us = (C.c_double * 7)()
capi.check(ctx.h, d.qb200_bench_update_variants(ctx.h, 2048, 63, us))
names = ["production (3M)", "no operand-sum DADDs", "no global stores", "no cp.async refill", "none of the three",
         "none of the three, no barrier", "4M product"]
print("jacobi_update_kernel at 2048 x 2048, microseconds per launch (ideal DMMA time at 37.1 TFLOP/s: 3M 43.4, 4M 57.9)")
for n, v in zip(names, us):
    print(f"{n:40s} {v:8.2f}")
