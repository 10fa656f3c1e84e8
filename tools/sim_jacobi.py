"""CPU simulation of the block one-sided Jacobi iteration of csrc/svd_jacobi.cu: how many sweeps do variants of the
pivot strategy / preconditioner need?  (NumPy only; design aid, not part of the product.)"""
import sys
import time
import numpy as np

rng = np.random.default_rng(7)


def crand(*s):
    return rng.standard_normal(s) + 1j * rng.standard_normal(s)


def rr_pairs(n, step):
    out = []
    for k in range(n // 2):
        if k == 0:
            a, b = n - 1, step
        else:
            a, b = (step + k) % (n - 1), (step - k + (n - 1)) % (n - 1)
        out.append((min(a, b), max(a, b)))
    return out


def theta_like(chi, decay=4.0):
    """(chi*2) x (2*chi) two-site wave function with Schmidt grading on all three bonds and a Haar gate applied."""
    lam = np.exp(-np.linspace(0, decay, chi))
    lam /= np.linalg.norm(lam)
    a = crand(chi, 2, chi) / np.sqrt(chi)
    b = crand(chi, 2, chi) / np.sqrt(chi)
    th = np.einsum("l,lam,m,mbr,r->labr", lam, a, lam, b, lam, optimize=True)
    g, _ = np.linalg.qr(crand(4, 4))
    th = np.einsum("xyab,labr->lxyr", g.reshape(2, 2, 2, 2), th, optimize=True)
    return th.reshape(2 * chi, 2 * chi)


def precondition(a, nqr=1):
    idx = np.argsort(-np.linalg.norm(a, axis=0), kind="stable")
    r = np.linalg.qr(a[:, idx], mode="r")
    x = r.conj().T
    for _ in range(nqr - 1):
        r = np.linalg.qr(x, mode="r")
        x = r.conj().T
    return np.ascontiguousarray(x)


def inner_cross_sweep(g, jb):
    """one cyclic sweep of the jb*jb cross rotations on the 2jb x 2jb Gram g (two-sided); returns W."""
    n2 = 2 * jb
    w = np.eye(n2, dtype=complex)
    g = g.copy()
    for t in range(jb):
        p = np.arange(jb)
        q = jb + (p + t) % jb
        a = g[p, p].real
        b = g[q, q].real
        c = g[p, q]
        n = np.abs(c)
        ok = n > 1e-300
        zeta = np.where(ok, 0.5 * (b - a) / np.where(ok, n, 1), 0)
        tt = np.where(ok, np.copysign(1.0, zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta)), 0)
        cs = 1 / np.sqrt(1 + tt * tt)
        sn = np.where(ok, tt * cs * c / np.where(ok, n, 1), 0)
        j = np.eye(n2, dtype=complex)
        j[p, p] = cs
        j[q, q] = cs
        j[p, q] = sn
        j[q, p] = -np.conj(sn)
        g = j.conj().T @ g @ j
        w = w @ j
    return w


def jacobi(x, jb=32, pivot="cross1", max_sweeps=40, tol=1e-7, verbose=False):
    k = x.shape[1]
    nb = k // jb
    x = x.copy()
    hist = []
    for sweep in range(max_sweeps):
        worst = 0.0
        for step in range(nb - 1):
            for (i, j) in rr_pairs(nb, step):
                cols = np.r_[i * jb:(i + 1) * jb, j * jb:(j + 1) * jb]
                p = x[:, cols]
                g = p.conj().T @ p
                d = np.sqrt(np.abs(np.diag(g).real))
                d[d == 0] = 1
                c = np.abs(g) / d[:, None] / d[None, :]
                np.fill_diagonal(c, 0)
                if pivot == "cross1" and step > 0:
                    c[:jb, :jb] = 0
                    c[jb:, jb:] = 0
                worst = max(worst, c.max())
                if pivot == "cross1":
                    if step == 0:
                        ev, w = np.linalg.eigh(g)  # stand-in for the full inner sweep at step 0
                        w = w[:, ::-1]
                    else:
                        w = inner_cross_sweep(g, jb)
                elif pivot == "full":
                    ev, w = np.linalg.eigh(g)
                    w = w[:, ::-1]
                x[:, cols] = p @ w
        hist.append(worst)
        if verbose:
            print(f"   sweep {sweep} worst {worst:.3e}", flush=True)
        if worst <= tol:
            break
    return len(hist), hist, x


if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    cases = {"ginibre": crand(k, k), "theta(decay 4)": theta_like(k // 2, 4.0), "theta(decay 12)": theta_like(k // 2, 12.0)}
    for name, a in cases.items():
        s_ref = np.linalg.svd(a, compute_uv=False)
        for nqr in (1, 2):
            x0 = precondition(a, nqr)
            for pivot, jb in (("cross1", 32), ("full", 32), ("full", 64), ("full", 128)):
                if k // jb < 4:
                    continue
                t0 = time.time()
                ns, hist, x = jacobi(x0, jb, pivot)
                s = np.sort(np.linalg.norm(x, axis=0))[::-1]
                print(f"{name:16s} k={k} qr x{nqr} pivot={pivot:6s} jb={jb:3d}: sweeps {ns:2d}  "
                      f"err {np.abs(s - s_ref).max() / s_ref[0]:.1e}  hist {' '.join('%.0e' % h for h in hist)}  ({time.time() - t0:.0f}s)",
                      flush=True)
