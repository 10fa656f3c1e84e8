import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
rng = np.random.default_rng(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
a = ctx.array(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
q, r = qb.qr(a, (0, 1), 1)
ctx.synchronize()
