"""Experiment: Jacobi sweeps on A vs on R^H (QR preconditioning) for random and TEBD-like matrices."""
import sys
import numpy as np
sys.path.insert(0, ".")
import qrochet_b200 as qb
ctx = qb.Context(0)
rng = np.random.default_rng(0)
def crand(*s): return rng.standard_normal(s) + 1j * rng.standard_normal(s)
def run(name, a):
    ctx.timer_begin(); qb.svd(ctx.array(a), (0, 1), 1); ms = ctx.timer_end()
    print(f"{name:40s} {a.shape} sweeps {ctx.svd_last_sweeps():3d}  {ms:8.1f} ms", flush=True)
for n in (512, 1024):
    a = crand(n, n)
    run("random A", a)
    q, r = np.linalg.qr(a)
    run("random R^H", r.conj().T.copy())
    q2, r2 = np.linalg.qr(r.conj().T)
    run("random R2^H (two QRs)", r2.conj().T.copy())
    # sorted columns by norm then QR
    # TEBD-like: theta = Dl Y Dr with decaying Schmidt spectra
    lam = np.exp(-np.linspace(0, 6, n // 2)); lam /= np.linalg.norm(lam)
    dl = np.repeat(lam, 2); dr = np.repeat(lam, 2)
    y, _ = np.linalg.qr(crand(n, n))
    th = (dl[:, None] * crand(n, n)) * dr[None, :]
    run("theta-like A (graded e^-6)", th)
    run("theta-like A^H", th.conj().T.copy())
    q, r = np.linalg.qr(th)
    run("theta-like R^H", r.conj().T.copy())
    idx = np.argsort(-np.linalg.norm(th, axis=0))
    q, r = np.linalg.qr(th[:, idx])
    run("theta-like R^H (cols sorted)", r.conj().T.copy())
