"""Experiment for the round-2 plan: how many FP64 Jacobi sweeps remain when the iteration starts from
X0 V with V the right singular vectors of a LOWER-PRECISION SVD (orthonormalised in FP64)?  The low-precision SVD is
LAPACK cgesdd on the host here (a stand-in for an FP32/TF32 Jacobi on the tensor path); `noise` degrades V further to
mimic plain-TF32 accuracy.  Prints the sweep counts of qb.svd on A and on A V."""
import json
import sys
import time
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(5)
def crand(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
def sweeps(a):
    d = ctx.array(np.asfortranarray(a))
    ctx.timer_begin()
    u, s, vc, kept, dw = qb.svd(d, (0, 1), 1)
    ms = ctx.timer_end()
    return ctx.svd_last_sweeps(), ms, s.to_host()
out = {"n": n}
cases = {"ginibre": crand(n, n)}
# TEBD-like: theta = (Λ Γ Λ)(Γ Λ) with a Haar gate folded in is what the bench factorises; a product of two random
# matrices with a graded diagonal in between has the same character (graded rows / columns, flat bulk)
lam = np.sort(rng.random(n // 2))[::-1] + 0.05
g = crand(n, n // 2) * lam[None, :]
cases["graded product"] = g @ crand(n // 2, n) + 1e-3 * crand(n, n)
for name, a in cases.items():
    t0 = time.time()
    u32, s32, vh32 = np.linalg.svd(a.astype(np.complex64), full_matrices=False)
    t_svd32 = time.time() - t0
    res = {"host_cgesdd_s": round(t_svd32, 2)}
    k0, ms0, s_ref = sweeps(a)
    res["fp64_jacobi_alone"] = {"sweeps": k0, "ms": round(ms0, 1)}
    for noise in (0.0, 1e-5, 1e-3):
        v = vh32.conj().T.astype(np.complex128)
        if noise:
            v = v + noise * crand(n, n) / np.sqrt(n)
        q, r = np.linalg.qr(v)  # exactly unitary in FP64
        b = a @ q
        k1, ms1, s1 = sweeps(b)
        res[f"after_lowp_V_noise_{noise:g}"] = {"sweeps": k1, "ms": round(ms1, 1),
                                               "sigma_err": float(np.abs(s1 - s_ref).max() / s_ref[0])}
    out[name] = res
    print(name, json.dumps(res), flush=True)
json.dump(out, open("gpurun_out/exp_mixed_precond.json", "w"), indent=1)
