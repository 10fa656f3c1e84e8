"""Step-by-step canonize! of an MPO-applied state (n = 20, chi = 512: first size that breaks) on the label-driven chain."""
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
from qrochet_b200.chain import Chain, site
from oracle import chain as oc
ctx = qb.Context(0)
n, chi = 20, 512
arrays = qb.rand_mps_arrays(np.random.default_rng(1003), n, chi)
mpo = qb.heisenberg_mpo_arrays(n)
harr = oc.apply_mpo_arrays(arrays, mpo)
g = qb.B200MPS(ctx, harr)
n0 = g.norm()
print("norm", n0, flush=True)
# factorisations of the widest site, checked on the host
s = n // 2
a = g.site(s)                                           # (l, o, r) = (2560, 2, 2560)
m_left = np.asfortranarray(a.reshape(-1, a.shape[2], order="F"))             # (l o) x r, what the SVD sweep factors
m_right = np.asfortranarray(a.reshape(a.shape[0], -1, order="F").conj().T)   # (o r) x l, what the QR sweep factors
for name, m in (("qr-sweep matrix", m_right), ("svd-sweep matrix", m_left)):
    sv = np.linalg.svd(m, compute_uv=False)
    rank = int(np.sum(sv > 1e-12 * sv[0]))
    t = ctx.array(m)
    q, r = qb.qr(t, [0, 1], 1)
    qh, rh = q.to_host(), r.to_host()
    e = qh.conj().T @ qh - np.eye(qh.shape[1])
    dr = np.abs(np.diag(rh))
    good = dr > 1e-10 * dr.max()
    print(name, m.shape, "rank", rank, "|QR-A|/|A|", np.linalg.norm(qh @ rh - m) / np.linalg.norm(m), "good cols", int(good.sum()),
          "|QhQ-I| good-good", np.abs(e[np.ix_(good, good)]).max(), "good-null", np.abs(e[np.ix_(good, ~good)]).max() if (~good).any() else 0,
          "null-null", np.abs(e[np.ix_(~good, ~good)]).max() if (~good).any() else 0, flush=True)
    u, sg, vc, kept, dw = qb.svd(t, [0, 1], 1)
    uh, sh, vh = u.to_host(), sg.to_host(), vc.to_host()
    k = len(sh)
    rec = (uh[:, :k] * sh[None, :]) @ vh[:, :k].T
    gs = sh > 1e-10 * sh[0]
    eu = uh.conj().T @ uh - np.eye(k)
    ev = vh.conj().T @ vh - np.eye(k)
    print("   svd: |USVt-A|/|A|", np.linalg.norm(rec - m) / np.linalg.norm(m), "sigma ok", np.abs(sh - sv[:k]).max() / sv[0],
          "|UhU-I| good", np.abs(eu[np.ix_(gs, gs)]).max(), "all", np.abs(eu).max(), "|VhV-I| good", np.abs(ev[np.ix_(gs, gs)]).max(),
          "all", np.abs(ev).max(), "noise sigma max", sh[~gs].max() if (~gs).any() else 0, flush=True)
    del t, q, r, u, sg, vc
# the sweeps, one site at a time
c = Chain(ctx, harr)
ref = Chain(ctx, harr)
for i in range(n, 1, -1):
    c.canonize_site(site(i), "left", "qr")
    if 6 <= i <= 15:
        t = c.tensor_at(site(i))
        print("  qr site", i, "tensor", t.data.shape, "norm", c.norm(), "overlap/n0^2", c.overlap(ref) / n0 ** 2, flush=True)
ov = c.overlap(ref)
print("after the QR sweep: norm", c.norm(), "overlap/n0^2", ov / n0 ** 2, flush=True)
lams = []
from qrochet_b200.chain import contract
for i in range(1, n):
    c.canonize_site(site(i), "right", "svd")
    lam = c.lambda_between(site(i), site(i + 1))
    print("  svd site", i, "norm", c.norm(), "lambda head", lam.to_host()[:2], "tail", lam.to_host()[-1], flush=True)
    c.tn.pop(lam)
    a = c.tensor_at(site(i + 1))
    c.tn.replace_tensor(a, contract(a, lam, dims=()))
    lams.append(lam)
print("after the SVD sweep (A form): norm", c.norm(), "overlap/n0^2", c.overlap(ref) / n0 ** 2, flush=True)
