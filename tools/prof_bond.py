"""One bulk TEBD bond update at chi=1024 (theta 2048 x 2048) for ncu captures."""
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
psi = qb.B200MPS.from_sites(ctx, sites, lams, form=1)
kept, dw = psi.evolve(qb.haar_gate(rng), [2, 3], maxdim=chi, iscanonical=True, renormalize=True)
ctx.synchronize()
print("kept", kept, "sweeps", ctx.svd_last_sweeps())
