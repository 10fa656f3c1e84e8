"""DMMA issue-pattern micro-benchmark (qb200_bench_dmma_patterns): what a complex-multiply DMMA stream can reach."""
import ctypes as C
import sys
sys.path.insert(0, ".")
import qrochet_b200 as qb
from qrochet_b200 import _capi as capi
ctx = qb.Context(0)
out = (C.c_double * 9)()
capi.check(ctx.h, capi.lib.qb200_bench_dmma_patterns(ctx.h, out))
names = ["independent acc, fixed operands", "complex pattern, register operands", "complex pattern, A from LDS.128"]
print("TFLOP/s at 8 / 16 / 32 warps per SM")
for m in range(3):
    print(f"{names[m]:40s}", " ".join(f"{out[m * 3 + o]:7.2f}" for o in range(3)))
