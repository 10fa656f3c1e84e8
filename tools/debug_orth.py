import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
rng = np.random.default_rng(5)
for shape in [(256, 256), (512, 512), (300, 129), (1024, 1024)]:
    a = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    u, s, vc, kept, dw = qb.svd(ctx.array(a), (0, 1), 1)
    u, s, vc = u.to_host(), s.to_host(), vc.to_host()
    k = min(shape)
    q, r = qb.qr(ctx.array(a), (0, 1), 1)
    q = q.to_host()
    print(shape, "U orth %.2e V orth %.2e rec %.2e  QR orth %.2e" % (
        np.abs(u.conj().T @ u - np.eye(k)).max(), np.abs(vc.T @ vc.conj() - np.eye(k)).max(),
        np.abs((u * s) @ vc.T - a).max() / s[0], np.abs(q.conj().T @ q - np.eye(k)).max()))
