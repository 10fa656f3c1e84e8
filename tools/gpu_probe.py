"""First-contact probe on a B200: DMMA peak, GEMM and SVD timings at the bench shapes."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import qrochet_b200 as qb

ctx = qb.Context(0)
out = {"dmma_peak_tflops": ctx.dmma_peak_tflops()}
print(out, flush=True)
rng = np.random.default_rng(0)


def crand(*s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


for (m, n, k) in [(2048, 2048, 1024), (4096, 4096, 4096)]:
    a, b = ctx.array(crand(m, k)), ctx.array(crand(k, n))
    c = ctx.empty((m, n))
    qb.contract(a, (0, 1), b, (1, 2), (0, 2), out=c)
    ctx.timer_begin()
    for _ in range(5):
        qb.contract(a, (0, 1), b, (1, 2), (0, 2), out=c)
    ms = ctx.timer_end() / 5
    out[f"gemm_{m}x{n}x{k}_ms"] = ms
    out[f"gemm_{m}x{n}x{k}_tflops"] = 8.0 * m * n * k / ms / 1e9
    print(m, n, k, ms, 8.0 * m * n * k / ms / 1e9, flush=True)

for n in [256, 512, 1024, 2048]:
    a = ctx.array(crand(n, n))
    t0 = time.perf_counter()
    ctx.timer_begin()
    u, s, vc, kept, dw = qb.svd(a, (0, 1), 1)
    ms = ctx.timer_end()
    out[f"svd_{n}_ms"] = ms
    out[f"svd_{n}_sweeps"] = ctx.svd_last_sweeps()
    out[f"svd_{n}_alg_tflops"] = 4 * (14 + 8) * n ** 3 / ms / 1e9
    print("svd", n, ms, ctx.svd_last_sweeps(), time.perf_counter() - t0, flush=True)

for (m, n) in [(2048, 1024), (1024, 512)]:
    a = ctx.array(crand(m, n))
    qb.qr(a, (0, 1), 1)
    ctx.timer_begin()
    qb.qr(a, (0, 1), 1)
    ms = ctx.timer_end()
    out[f"qr_{m}x{n}_ms"] = ms
    print("qr", m, n, ms, flush=True)
json.dump(out, open("gpurun_out/probe.json", "w"), indent=1)
