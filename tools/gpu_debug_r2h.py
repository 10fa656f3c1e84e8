"""Where does canonize!/compress of an MPO-applied state (rank-deficient site tensors) stop preserving the state?"""
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
for n, chi in ((16, 512), (18, 512), (20, 512), (24, 512)):
    arrays = qb.rand_mps_arrays(np.random.default_rng(1003), n, chi)
    mpo = qb.heisenberg_mpo_arrays(n)
    g = qb.B200MPS(ctx, arrays).apply_mpo(mpo)
    n0 = g.norm()
    c = g.copy().canonize()
    ov = c.overlap(g)
    print(n, chi, "max bond", max(g.bond_dims()), "norm before", n0, "after canonize!", c.norm(), "|<c|g>|/n0^2", abs(ov) / n0 ** 2,
          "bond-1 lambda", c.lambdas()[0], flush=True)
    # single QR steps on the widest site, checked on the host
    s = n // 2
    a = g.site(s)
    m = np.asfortranarray(a.reshape(a.shape[0], -1, order="F").conj().T)   # (o r) x l, what right_canonize_qr factors
    t = ctx.array(m)
    q, r = qb.qr(t, [0, 1], 1)
    qh, rh = q.to_host(), r.to_host()
    print("   QR", m.shape, "|QR - A|/|A|", np.linalg.norm(qh @ rh - m) / np.linalg.norm(m), flush=True)
    del t, q, r
