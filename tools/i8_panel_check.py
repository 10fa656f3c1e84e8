"""First run of the INT8 (Ozaki) panel product next round: qb200_i8_panel_gemm against its bit-exact specification
(tools/exp_ozaki.py::ozaki_complex) and against the FP64 product, then a timing against the DMMA GEMM of the same shape."""
import ctypes as C
import sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import qrochet_b200 as qb
from qrochet_b200 import _capi as capi
import exp_ozaki

ctx = qb.Context(0)
rng = np.random.default_rng(0)
for m, grade in ((128, None), (1000, None), (2048, np.logspace(0, -8, 64))):
    x = rng.standard_normal((m, 64)) + 1j * rng.standard_normal((m, 64))
    if grade is not None:
        x = x * grade[None, :]
    w = np.linalg.qr(rng.standard_normal((64, 64)) + 1j * rng.standard_normal((64, 64)))[0]
    a, b = ctx.array(x), ctx.array(w)
    c = ctx.empty((m, 64))
    code = capi.lib.qb200_i8_panel_gemm(ctx.h, a.h, b.h, c.h)
    if code != 0:
        print("error", code, capi.lib.qb200_last_error(ctx.h))
        sys.exit(1)
    got = c.to_host()
    spec = exp_ozaki.ozaki_complex(x, w, 8)
    want = x @ w
    col = lambda d: float(np.max(np.linalg.norm(d, axis=0) / np.linalg.norm(want, axis=0)))
    print(f"M = {m:5d}{' graded' if grade is not None else '       '}: bit-exact vs specification: {np.array_equal(got, spec)}"
          f" (max |diff| {np.abs(got - spec).max():.2e}); column-wise error vs FP64 product {col(got - want):.2e}")
m = 2048 * 16
x = rng.standard_normal((m, 64)) + 1j * rng.standard_normal((m, 64))
a, b, c = ctx.array(x), ctx.array(w), ctx.empty((m, 64))
for name, fn in (("INT8 Ozaki panel product", lambda: capi.lib.qb200_i8_panel_gemm(ctx.h, a.h, b.h, c.h)),
                 ("DMMA GEMM (3M)", lambda: qb.contract(a, (0, 1), b, (1, 2), (0, 2), out=c))):
    fn(); ctx.synchronize()
    best = 1e9
    for _ in range(3):
        ctx.timer_begin(); fn(); best = min(best, ctx.timer_end())
    print(f"{name:26s} {m} x 64 x 64: {best:.3f} ms = {8.0 * m * 64 * 64 / best / 1e9:.1f} TFLOP/s (algorithmic)")
