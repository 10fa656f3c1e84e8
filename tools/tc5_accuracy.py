"""Where does the ComplexF32 tcgen05 GEMM lose accuracy?  Same products through the mma.sync kernel (default) and the
tcgen05 kernel (QB200_C64_TCGEN05=1, set by the caller), with operands that are exactly TF32-representable (lo = 0: only
the hh term is non-zero, so the error is the tensor core's own product / accumulation error) and with full FP32 operands."""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
rng = np.random.default_rng(3)
def crand32(*s): return (rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64)
def trunc_tf32(x):
    v = x.copy().view(np.uint32)
    v &= np.uint32(0xffffe000)
    return v.view(np.complex64)
print("kernel:", "tcgen05" if os.environ.get("QB200_C64_TCGEN05") == "1" else "mma.sync")
for k in (64, 512, 4096):
    a, b = crand32(256, k), crand32(k, 192)
    for name, (x, y) in {"fp32 operands": (a, b), "tf32-exact operands": (trunc_tf32(a), trunc_tf32(b)),
                         "positive tf32-exact": (trunc_tf32(np.abs(a.real).astype(np.complex64)), trunc_tf32(np.abs(b.real).astype(np.complex64)))}.items():
        got = qb.contract(ctx.array(x), (0, 1), ctx.array(y), (1, 2), (0, 2)).to_host()
        want = x.astype(np.complex128) @ y.astype(np.complex128)
        d = got - want
        print(f"K={k:5d} {name:22s} max rel err {np.abs(d).max() / np.abs(want).max():.3e}  mean signed re err / max {d.real.mean() / np.abs(want).max():+.3e}")
