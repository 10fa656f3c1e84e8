"""Phase profile of config 3's compression (n sites, chi, Heisenberg MPO D = 5: apply_mpo then compress(maxdim=chi)):
python tools/prof_compress.py [n] [chi]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
chi = int(sys.argv[2]) if len(sys.argv) > 2 else 512
ctx = qb.Context(0)
psi = qb.B200MPS(ctx, qb.rand_mps_arrays(np.random.default_rng(1003), n, chi))
mpo = qb.heisenberg_mpo_arrays(n)
phi = psi.copy()
phi.apply_mpo(mpo)
ctx.synchronize()
ctx.profile(True); ctx.profile_read()
t0 = time.time()
phi.compress(maxdim=chi)
ctx.synchronize()
wall = time.time() - t0
pp = ctx.profile_read(); ctx.profile(False)
print(json.dumps({"n": n, "chi": chi, "wall_s": wall, "svd_totals": list(ctx.svd_totals()),
                  "phases": {k: [v[0], round(v[1], 2)] for k, v in pp.items() if v[0]}}))
