import sys, time
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
n, chi = int(sys.argv[1]), int(sys.argv[2])
ctx = qb.Context(0)
arrays = qb.rand_mps_arrays(np.random.default_rng(1004), n, chi)
psi = qb.B200MPS(ctx, arrays)
t0 = time.time()
try:
    psi.canonize()
except Exception as e:
    print("FAILED", e)
print("canonize", time.time() - t0)
