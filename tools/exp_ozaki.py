"""CPU experiment for the round-2 plan: FP64 products on the INT8 tensor pipe by Ozaki splitting, for
the Jacobi update X_p <- X_p W_p (K = 64).  Emulates exactly what the INT8 kernel would compute (7-bit signed slices,
row-wise shared exponents for X, column-wise for W, exact integer accumulation of the slice products with
i + j < nslices) and measures the COLUMN-WISE relative error the one-sided Jacobi needs, for flat and for graded column
norms, against the DMMA-style floating-point product; then runs a small one-sided block Jacobi with the emulated update
and compares the singular values with LAPACK."""
import sys
import numpy as np

BITS = 7


def split_rows(a, ns):
    """a (m x k) real -> integer slices s[0..ns) (int64 holding 7-bit signed values) and row scales: a ~ 2^e_r * sum_t
    s_t 2^(-BITS (t+1))"""
    amax = np.abs(a).max(axis=1)
    e = np.where(amax > 0, np.ceil(np.log2(np.maximum(amax, 1e-300))) + 1, 0.0)   # |a| / 2^e < 1/2
    r = a / np.exp2(e)[:, None]
    slices = []
    for _ in range(ns):
        r = r * (1 << BITS)
        s = np.rint(r)          # |s| <= 64: fits a signed 8-bit slice; the remainder is in [-1/2, 1/2]
        slices.append(s.astype(np.int64))
        r = r - s
    return slices, e


def ozaki_real(a, b, ns):
    """a (m x k) @ b (k x n) with ns slices each, products with i + j < ns (what ns (ns + 1) / 2 INT8 GEMMs give)"""
    sa, ea = split_rows(a, ns)
    sb, eb = split_rows(b.T.copy(), ns)
    out = np.zeros((a.shape[0], b.shape[1]))
    for i in range(ns):
        for j in range(ns - i):
            p = sa[i] @ sb[j].T      # exact: |entries| <= 64 * 64 * k fits int32 for k = 64
            out += p.astype(np.float64) * np.exp2(-BITS * (i + j + 2))
    return out * np.exp2(ea)[:, None] * np.exp2(eb)[None, :]


def ozaki_complex(x, w, ns):
    re = ozaki_real(x.real, w.real, ns) - ozaki_real(x.imag, w.imag, ns)
    im = ozaki_real(x.real, w.imag, ns) + ozaki_real(x.imag, w.real, ns)
    return re + 1j * im


def colwise_err(got, want):
    return float(np.max(np.linalg.norm(got - want, axis=0) / np.linalg.norm(want, axis=0)))


def haar(n, rng):
    q, r = np.linalg.qr(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    return q * (np.diag(r) / np.abs(np.diag(r)))


def main():
    rng = np.random.default_rng(0)
    m, k = 512, 64
    print("column-wise relative error of X W (512 x 64 times 64 x 64 unitary), exact reference in longdouble")
    for name, grade in (("flat", np.ones(k)), ("graded 1e-8", np.logspace(0, -8, k)), ("graded 1e-14", np.logspace(0, -14, k))):
        x = (rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k))) * grade[None, :]
        for wname, w in (("Haar W", haar(k, rng)), ("W = I + 1e-6 skew", None)):
            if w is None:
                s = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
                w = np.eye(k) + 1e-6 * (s - s.conj().T)
                w = np.linalg.qr(w)[0]
            want = (x.astype(np.clongdouble) @ w.astype(np.clongdouble)).astype(np.complex128)
            line = f"  {name:13s} {wname:18s} fp64 {colwise_err(x @ w, want):.1e}"
            for ns in (6, 7, 8, 9, 10):
                line += f" | {ns} slices {colwise_err(ozaki_complex(x, w, ns), want):.1e}"
            print(line, flush=True)
    # column scaling fixes the graded case: X = Xn D, X W = Xn (D W D'^-1) D' with D' = the new column norms is not
    # available in advance, but scaling X's columns to unit norm and W's rows accordingly keeps every slice meaningful:
    print("with the columns of X scaled to unit norm before slicing (X = Xn D; X W = Xn (D W)):")
    for name, grade in (("graded 1e-8", np.logspace(0, -8, k)), ("graded 1e-14", np.logspace(0, -14, k))):
        x = (rng.standard_normal((m, k)) + 1j * rng.standard_normal((m, k))) * grade[None, :]
        s = rng.standard_normal((k, k)) + 1j * rng.standard_normal((k, k))
        w = np.linalg.qr(np.eye(k) + 1e-6 * (s - s.conj().T))[0]
        want = (x.astype(np.clongdouble) @ w.astype(np.clongdouble)).astype(np.complex128)
        d = np.linalg.norm(x, axis=0)
        line = f"  {name:13s} W = I + 1e-6 skew  fp64 {colwise_err(x @ w, want):.1e}"
        for ns in (8, 9, 10, 12):
            line += f" | {ns} slices {colwise_err(ozaki_complex(x / d[None, :], d[:, None] * w, ns), want):.1e}"
        print(line, flush=True)

    # small one-sided block Jacobi (blocks of 32, pairs of 64 columns) with the emulated update
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    s_ref = np.linalg.svd(a, compute_uv=False)
    for ns in (None, 8, 9):
        x = np.linalg.qr(a)[1].conj().T.copy()
        nb = n // 32
        for sweep in range(30):
            worst = 0.0
            for step in range(nb - 1):
                for pr in range(nb // 2):
                    if pr == 0:
                        p_, q_ = nb - 1, step
                    else:
                        p_, q_ = (step + pr) % (nb - 1), (step - pr) % (nb - 1)
                    p_, q_ = min(p_, q_), max(p_, q_)
                    cols = np.r_[p_ * 32:(p_ + 1) * 32, q_ * 32:(q_ + 1) * 32]
                    panel = x[:, cols]
                    g = panel.conj().T @ panel
                    dn = np.sqrt(np.abs(np.diag(g)))
                    off = np.abs(g - np.diag(np.diag(g))) / np.outer(dn, dn)
                    worst = max(worst, off.max())
                    _, wv = np.linalg.eigh(g)
                    wv = wv[:, ::-1]
                    x[:, cols] = panel @ wv if ns is None else ozaki_complex(panel, wv, ns)
            if worst < 1e-7:
                break
        s = np.sort(np.linalg.norm(x, axis=0))[::-1]
        print(f"block Jacobi n = {n}, update {'fp64' if ns is None else str(ns) + ' int8 slices'}: {sweep + 1} sweeps, "
              f"max |sigma - LAPACK| / sigma_1 = {np.abs(s - s_ref).max() / s_ref[0]:.2e}", flush=True)


if __name__ == "__main__":
    main()
