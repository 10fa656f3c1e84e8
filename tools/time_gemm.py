"""K1 timing: plain and permuted ComplexF64 GEMMs (CUDA events), TFLOP/s = 8 M N K / t."""
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
rng = np.random.default_rng(0)
def crand(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
def run(name, a, ma, b, mb, mc, flops, reps=3):
    A, B = ctx.array(a), ctx.array(b)
    qb.contract(A, ma, B, mb, mc)
    ctx.synchronize()
    best = 1e9
    for _ in range(reps):
        ctx.timer_begin(); qb.contract(A, ma, B, mb, mc); best = min(best, ctx.timer_end())
    print(f"{name:46s} {best:8.3f} ms {flops / best / 1e9:7.2f} TF/s", flush=True)
for n in (2048, 4096):
    a, b = crand(n, n), crand(n, n)
    run(f"plain {n}^3", a, (0, 1), b, (1, 2), (0, 2), 8.0 * n ** 3)
    run(f"A^T (k fastest) {n}^3", a, (1, 0), b, (1, 2), (0, 2), 8.0 * n ** 3)
# QR trailing-update shapes: C = Q_p^H T (k = 2048 rows, 128 x 1920), T -= Q_p C
q, t = crand(2048, 128), crand(2048, 1920)
run("Qp^H T: 128 x 1920 x 2048", q, (1, 0), t, (1, 2), (0, 2), 8.0 * 128 * 1920 * 2048)
c = crand(128, 1920)
run("Qp C: 2048 x 1920 x 128", q, (0, 1), c, (1, 2), (0, 2), 8.0 * 2048 * 1920 * 128)
# theta GEMM of a bulk bond and a 12-mode permuted contraction like the sliced tree's top node
run("theta 2048 x 2048 x 1024", crand(2048, 1024), (0, 1), crand(1024, 2048), (1, 2), (0, 2), 8.0 * 2048 * 2048 * 1024)
sh = (4,) * 6
x, y = crand(*(sh + sh)), crand(*(sh + sh))   # 4096 x 4096 each as 12 modes
ma = tuple(range(12)); mb = (6, 13, 7, 14, 8, 15, 9, 16, 10, 17, 11, 18); mc = (0, 13, 1, 14, 2, 15, 3, 16, 4, 17, 5, 18)
run("12-mode interleaved 4096^3", x, ma, y, mb, mc, 8.0 * 4096 ** 3)
