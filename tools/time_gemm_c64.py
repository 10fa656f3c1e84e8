"""K1 (ComplexF32) timing: plain and permuted GEMMs on the 3xTF32 mma.sync path (CUDA events).
TFLOP/s = 8 M N K / t (algorithmic complex flops); executed HMMA flops are 3x that (three TF32 products per real
product), so the tensor-pipe fraction is 3 * TF/s / (TF32 mma.sync peak)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
tf32_peak, bf16_peak = ctx.hmma_peak_tflops()
print(f"mma.sync peaks: TF32 {tf32_peak:.1f} TF/s, BF16 {bf16_peak:.1f} TF/s")
rng = np.random.default_rng(0)
def crand(*s): return (rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64)
def run(name, a, ma, b, mb, mc, flops, reps=3):
    A, B = ctx.array(a), ctx.array(b)
    qb.contract(A, ma, B, mb, mc)
    ctx.synchronize()
    best = 1e9
    for _ in range(reps):
        ctx.timer_begin(); qb.contract(A, ma, B, mb, mc); best = min(best, ctx.timer_end())
    tf = flops / best / 1e9
    print(f"{name:40s} {best:8.3f} ms {tf:7.2f} TF/s  HMMA pipe {3 * tf / tf32_peak:5.1%}", flush=True)
for n in (2048, 4096):
    a, b = crand(n, n), crand(n, n)
    run(f"plain {n}^3", a, (0, 1), b, (1, 2), (0, 2), 8.0 * n ** 3)
    run(f"A^T (k fastest) {n}^3", a, (1, 0), b, (1, 2), (0, 2), 8.0 * n ** 3)
run("theta 2048 x 2048 x 1024", crand(2048, 1024), (0, 1), crand(1024, 2048), (1, 2), (0, 2), 8.0 * 2048 * 2048 * 1024)
sh = (4,) * 6
x, y = crand(*(sh + sh)), crand(*(sh + sh))
ma = tuple(range(12)); mb = (6, 13, 7, 14, 8, 15, 9, 16, 10, 17, 11, 18); mc = (0, 13, 1, 14, 2, 15, 3, 16, 4, 17, 5, 18)
run("12-mode interleaved 4096^3", x, ma, y, mb, mc, 8.0 * 4096 ** 3)
