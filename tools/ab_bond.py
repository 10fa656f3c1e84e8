"""A/B timing of one bulk TEBD bond update (theta 2048 x 2048 at chi=1024) with the phase profiler on: prints the
per-phase milliseconds as JSON.  Run once per setting of QB200_UPDATE_3M / QB200_GEMM_3M (read at first use)."""
import json
import os
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = qb.Context(0)
rng = np.random.default_rng(0)
def crand(*s): return (rng.standard_normal(s) + 1j * rng.standard_normal(s)) / np.sqrt(s[0] * s[1])
sites = [np.asfortranarray(crand(1, 2, chi)), np.asfortranarray(crand(chi, 2, chi)), np.asfortranarray(crand(chi, 2, chi)),
         np.asfortranarray(crand(chi, 2, 1))]
lams = []
for _ in range(3):
    l = np.sort(rng.random(chi))[::-1] + 0.1
    lams.append(l / np.linalg.norm(l))
out = {"env": {k: os.environ.get(k) for k in ("QB200_UPDATE_3M", "QB200_GEMM_3M")}}
for rep in range(3):
    psi = qb.B200MPS.from_sites(ctx, sites, lams, form=1)
    ctx.profile(True)
    ctx.profile_read()
    kept, dw = psi.evolve(qb.haar_gate(np.random.default_rng(7)), [2, 3], maxdim=chi, iscanonical=True, renormalize=True)
    pp = ctx.profile_read()
    ctx.profile(False)
    lam = psi.lambdas()[1]
    out[f"rep{rep}"] = {"kept": int(kept), "dw": float(dw), "sweeps": ctx.svd_last_sweeps(),
                        "lam_head": [float(x) for x in lam[:3]], "lam_sum2": float(np.sum(lam ** 2)),
                        "phases_ms": {k: round(v[1], 3) for k, v in pp.items() if v[0]}}
    del psi
print(json.dumps(out))
