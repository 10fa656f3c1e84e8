import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
rng = np.random.default_rng(0)
n = 4096
a = ctx.array(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
b = ctx.array(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
c = ctx.empty((n, n))
for _ in range(3):
    qb.contract(a, (0, 1), b, (1, 2), (0, 2), out=c)
ctx.synchronize()
