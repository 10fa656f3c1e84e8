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
    """a (m x k) real -> ns integer slices (7-bit signed, |s| <= 64) and one exponent per row:
    a = 2^e_r * sum_t s_t 2^(-BITS (t + 1)) + remainder.  Every operation is exact in FP64, so the device kernel
    (i8_panel_gemm.cu) reproduces the slices bit for bit: e_r = frexp-exponent of the row maximum + 1 (0 for a zero
    row), r = a 2^-e_r (|r| < 1/2), then ns times: r <- 128 r, s = rint(r) (ties to even), r <- r - s."""
    amax = np.abs(a).max(axis=1)
    e = np.where(amax > 0, np.frexp(amax)[1] + 1, 0).astype(np.int64)
    r = np.ldexp(a, -e[:, None])
    slices = []
    for _ in range(ns):
        r = r * float(1 << BITS)
        s = np.rint(r)
        slices.append(s.astype(np.int64))
        r = r - s
    return slices, e


def ozaki_orders(a, b, ns):
    """integer accumulators of a (m x k) @ b (k x n), one per order d = i + j < ns (what TMEM holds), + exponents"""
    sa, ea = split_rows(a, ns)
    sb, eb = split_rows(b.T.copy(), ns)
    acc = [sum(sa[i] @ sb[d - i].T for i in range(d + 1)) for d in range(ns)]   # exact (int64 here, int32 on device)
    return acc, ea, eb


def ozaki_real(a, b, ns):
    acc, ea, eb = ozaki_orders(a, b, ns)
    out = np.zeros((a.shape[0], b.shape[1]))
    for d in range(ns):
        out += np.ldexp(acc[d].astype(np.float64), -BITS * (d + 2) + ea[:, None] + eb[None, :])
    return out


def ozaki_complex(x, w, ns):
    """3M product in the kernel's order of operations: products P = Xr Wr, Q = Xi Wi, S = (Xr + Xi)(Wr + Wi), one after
    the other, every order d added straight into Cr / Ci:  P: Cr += t, Ci -= t;  Q: Cr -= t, Ci -= t;  S: Ci += t."""
    cr = np.zeros((x.shape[0], w.shape[1]))
    ci = np.zeros_like(cr)
    for which, (a, b) in enumerate(((x.real, w.real), (x.imag, w.imag), (x.real + x.imag, w.real + w.imag))):
        acc, ea, eb = ozaki_orders(np.ascontiguousarray(a), np.ascontiguousarray(b), ns)
        for d in range(ns):
            t = np.ldexp(acc[d].astype(np.float64), -BITS * (d + 2) + ea[:, None] + eb[None, :])
            if which == 0:
                cr += t
                ci -= t
            elif which == 1:
                cr -= t
                ci -= t
            else:
                ci += t
    return cr + 1j * ci


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
