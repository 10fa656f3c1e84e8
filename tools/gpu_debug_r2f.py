"""Round-2 debugging aid: (1) the sliced executor on small networks (graph / no graph), (2) QR / compress on the
rank-deficient site tensors an MPO application produces (Cholesky-QR fast path vs Householder TSQR)."""
import os
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
from oracle import chain as oc
from oracle import circuit as ocirc

what = sys.argv[1]
ctx = qb.Context(0)
if what == "tn":
    for n, depth, target in ((10, 4, 2 ** 5), (12, 5, 2 ** 4), (14, 6, 2 ** 6), (16, 5, 0)):
        gates = qb.random_fsim_circuit(n, depth)
        ket, bra = ocirc.random_product_state(n, 1), ocirc.random_product_state(n, 2)
        arrays, modes = qb.amplitude_network(n, gates, ket, bra)
        want = ocirc.statevector_amplitude(n, gates, ket, bra)
        for opt in (1, 0):
            sc = qb.SlicedContraction(ctx, arrays, modes, target, optimizer=opt)
            got = sc.contract()
            got2 = sc.contract()
            parts = sum(sc.contract(r, 3) for r in range(3))
            print(n, depth, target, "opt", opt, "slices", sc.nslices, "err", abs(got - want), abs(got2 - want), abs(parts - want),
                  flush=True)
elif what == "qr":
    for n, chi in ((12, 64), (14, 128)):
        arrays = qb.rand_mps_arrays(np.random.default_rng(1003), n, chi)
        mpo = qb.heisenberg_mpo_arrays(n)
        g = qb.B200MPS(ctx, arrays).apply_mpo(mpo)
        ref = oc.Chain(oc.apply_mpo_arrays(arrays, mpo))
        n0 = g.norm()
        print(n, chi, "bond dims", g.bond_dims(), "norm device", n0, "oracle", ref.norm(), flush=True)
        # QR of the widest site as the QR sweep sees it: (p chi_r) x chi_l, i.e. the adjoint of the (l | o r) matricisation
        s = n // 2
        a = g.site(s)                                    # (l, o, r)
        m = a.reshape(a.shape[0], -1, order="F").conj().T   # (o r) x l
        t = ctx.array(np.asfortranarray(m))
        q, r = qb.qr(t, [0, 1], 1)
        qh, rh = q.to_host(), r.to_host()
        sv = np.linalg.svd(m, compute_uv=False)
        print("   site", s, "matrix", m.shape, "rank", int(np.sum(sv > 1e-12 * sv[0])), "|QR - A|/|A|",
              np.linalg.norm(qh @ rh - m) / np.linalg.norm(m), "|Q^H Q - I|", np.abs(qh.conj().T @ qh - np.eye(qh.shape[1])).max(),
              flush=True)
        c = g.copy().compress(maxdim=chi)
        o = oc.compress(ref.copy(), maxdim=chi)
        worst = max(np.abs(x - y).max() / y[0] for x, y in zip(c.lambdas(), o.lambdas()))
        print("   compress: worst dlambda/lambda_1", worst, "dims equal", c.bond_dims() == [len(l) for l in o.lambdas()], flush=True)
