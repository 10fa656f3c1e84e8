"""Wall times of BASELINE.json configs 1-3 on one B200 (secondary numbers of SURVEY.md §8d; the headline configs 4/5
are bench.py's).  Device-side times (CUDA events on the context stream), inputs already resident in HBM, one warm-up
run of each config first.  Prints one JSON object.

    python tools/run_configs.py [--skip3]
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import qrochet_b200 as qb  # noqa: E402

ctx = qb.Context(0)
out = {"device": "B200", "dtype": "ComplexF64", "timing": "CUDA events on the context stream, state resident in HBM"}


def timed(fn, reps=1):
    fn()  # warm-up (kernel attributes, workspace pool, tables)
    best = None
    for _ in range(reps):
        ctx.synchronize()
        ctx.timer_begin()
        r = fn()
        ms = ctx.timer_end()
        best = ms if best is None else min(best, ms)
    return best, r


# ---- config 1: rand MPS n=16 chi=32: canonize! + overlap + 1-site expect (Z on site 8) ----
arrays = qb.rand_mps_arrays(np.random.default_rng(1001), 16, 32)
base = qb.B200MPS(ctx, arrays)
Z = np.diag([1.0, -1.0]).astype(np.complex128)


def c1():
    psi = base.copy()
    psi.canonize()
    return psi.overlap(psi), psi.expect([Z], [8])[0]


ms, (ov, ez) = timed(c1, 3)
out["config1"] = {"workload": "n=16 chi=32: canonize! + overlap + <Z_8>", "ms": ms, "overlap": [ov.real, ov.imag],
                  "expect_Z8": [ez.real, ez.imag]}

# ---- config 2: n=64 chi=256, 4 brickwork layers of Haar gates, maxdim=256 ----
n, chi = 64, 256
psi2 = qb.B200MPS(ctx, qb.rand_mps_arrays(np.random.default_rng(1002), n, chi)).canonize()
odd, even = list(range(1, n, 2)), list(range(2, n, 2))
order = odd + even + odd + even
gates = [qb.haar_gate(np.random.default_rng(2000 + 64 * (i // 32) + b)) for i, b in enumerate(order)]


def c2():
    psi = psi2.copy()
    kept, dw = psi.evolve_circuit(gates, order, maxdim=chi, iscanonical=True, renormalize=True)
    return sum(kept), float(np.sum(dw)), psi.norm()


ms, (kept, dw, nrm) = timed(c2, 2)
out["config2"] = {"workload": "n=64 chi=256: 4 brickwork layers (126 evolve!), maxdim=256, renormalize", "ms": ms,
                  "kept_total": kept, "discarded_weight_total": dw, "norm_after": nrm,
                  "algorithmic_tflop": 2 * 0.587, "note": "0.587 TFLOP per 2-layer sweep (SURVEY §8d)"}

# ---- config 3: n=64 chi=512 x Heisenberg MPO D=5: mixed_canonize! + <H> + MPO application with truncation ----
if "--skip3" not in sys.argv:
    n, chi = 64, 512
    t0 = time.perf_counter()
    psi3 = qb.B200MPS(ctx, qb.rand_mps_arrays(np.random.default_rng(1003), n, chi))
    mpo = qb.heisenberg_mpo_arrays(n)
    setup = time.perf_counter() - t0
    res = {}

    def c3a():
        psi = psi3.copy()
        psi.mixed_canonize(32)
        return psi

    ms_mc, psi_mc = timed(c3a)
    ms_e, e = timed(lambda: psi3.expect_mpo(mpo))

    def c3c():
        phi = psi3.copy()
        phi.apply_mpo(mpo)
        return phi

    ms_ap, phi = timed(c3c)
    d_applied = max(phi.bond_dims())

    def c3d():
        w = phi.copy()
        w.compress(maxdim=chi)
        return w

    ms_cp, w = timed(c3d)
    # <H psi|H psi> before and after truncation: the discarded weight of the compression
    n2_full = phi.overlap(phi).real
    n2_trunc = w.overlap(w).real
    out["config3"] = {"workload": "n=64 chi=512, Heisenberg MPO D=5", "mixed_canonize_ms": ms_mc, "expect_mpo_ms": ms_e,
                      "expect_mpo": [e.real, e.imag], "apply_mpo_ms": ms_ap, "bond_after_apply": d_applied,
                      "compress_to_512_ms": ms_cp, "norm2_H_psi": n2_full, "norm2_after_truncation": n2_trunc,
                      "relative_discarded_weight": 1.0 - n2_trunc / n2_full, "host_setup_s": setup}
print(json.dumps(out))
