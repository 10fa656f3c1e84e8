"""Host-side input generators of the reference API: `rand(Chain, Open, State; n, χ, p, eltype)`
(/root/reference/src/Ansatz/Chain.jl:223-256) and Haar-random two-site gates.  Pure set-up code (K11 in
SURVEY.md §2.3): runs on the host, never inside a timed region."""
from __future__ import annotations

import numpy as np


def bond_dims(n: int, chi: int, p: int = 2):
    """Bond dimensions of `rand` (Chain.jl:230-236): bond b (1-based) has min(χ, p^b, p^(n-b))."""
    return [min(chi, p ** b, p ** (n - b)) for b in range(1, n)]


def rand_mps_arrays(rng: np.random.Generator, n: int, chi: int, p: int = 2):
    """Site arrays in the reference's default order (o, l, r): random row-orthonormal χl x (χr p) matrices
    (the reference orthonormalises the rows with Muscle.gramschmidt!; a Householder QR of the adjoint spans
    the same row space and is O(χ³) BLAS-3), reshaped (χl, χr, p) -> permuted (p, χl, χr); site 1 / sqrt(p)
    => right-canonical, norm 1."""
    arrays = []
    for i in range(1, n + 1):
        after_mid = i > n // 2
        j = (n + 1 - abs(2 * i - n - 1)) // 2
        chil, chir = min(chi, p ** (j - 1)), min(chi, p ** j)
        if n % 2 == 1 and i == n // 2 + 1:
            chir = chil
        elif after_mid:
            chil, chir = chir, chil
        if i == 1:
            chil, chir = chir, 1
        a = rng.random((chil, chir * p)) + 1j * rng.random((chil, chir * p))
        q, _ = np.linalg.qr(a.conj().T)
        a = np.reshape(q.conj().T, (chil, chir, p), order="F")
        arrays.append(np.transpose(a, (2, 0, 1)))
    arrays[0] = np.reshape(arrays[0], (p, p), order="F") / np.sqrt(p)
    arrays[-1] = np.reshape(arrays[-1], (p, p), order="F")
    return arrays


def rand_mpo_arrays(rng: np.random.Generator, n: int, chi: int, p: int = 2, complex_entries: bool = False):
    """`rand(Chain, Open, Operator; n, χ, p, eltype)` (Chain.jl:260-297): site arrays in the default MPO order
    (o, i, l, r), first (o, i, r), last (o, i, l).  Each site is a random row-orthonormal χl x (χr p²) matrix (QR of
    the adjoint, as in rand_mps_arrays), bond b = min(χ, p^(2b), p^(2(n-b))), site 1 / sqrt(min(χ, p²)) so that the
    operator has Frobenius norm 1.  Real entries by default like the reference (eltype = Float64, :264)."""
    arrays = []
    for i in range(1, n + 1):
        j = (n + 1 - abs(2 * i - n - 1)) // 2
        chil, chir = min(chi, p ** (2 * (j - 1))), min(chi, p ** (2 * j))
        if n % 2 == 1 and i == n // 2 + 1:
            chir = chil
        elif i > n // 2:
            chil, chir = chir, chil
        shape = (chir, p, p) if i == 1 else ((chil, p, p) if i == n else (chil, chir, p, p))
        cols = int(np.prod(shape[1:]))
        a = rng.random((shape[0], cols))
        if complex_entries:
            a = a + 1j * rng.random((shape[0], cols))
        q, _ = np.linalg.qr(a.conj().T)
        a = np.reshape(q.conj().T, shape, order="F")
        arrays.append(np.transpose(a, (1, 2, 0)) if i in (1, n) else np.transpose(a, (2, 3, 0, 1)))
    arrays[0] = arrays[0] / np.sqrt(min(chi, p * p))
    return arrays


def haar_gate(rng: np.random.Generator, d: int = 4):
    """Haar-random d x d unitary (QR of complex Ginibre, phases fixed) as the reference's gate array with
    dims (o1, o2, i1, i2), column-major reshape, first lane = fastest bit."""
    z = (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) / np.sqrt(2)
    q, r = np.linalg.qr(z)
    u = q * (np.diag(r) / np.abs(np.diag(r)))
    k = int(round(np.log2(d)))
    return np.reshape(u, (2,) * (2 * k), order="F")


def heisenberg_mpo_arrays(n: int, J: float = 1.0, h: float = 0.0):
    """Spin-1/2 Heisenberg chain H = J sum S_k.S_{k+1} + h sum Sz_k as an MPO with D = 5 in the reference's default
    MPO order (o, i, l, r) (Chain.jl:34); first site (o, i, r), last site (o, i, l) (BASELINE config 3)."""
    sz = np.diag([0.5, -0.5]).astype(np.complex128)
    sp = np.array([[0, 1], [0, 0]], dtype=np.complex128)
    sm = sp.T.copy()
    one = np.eye(2, dtype=np.complex128)
    w = np.zeros((5, 5, 2, 2), dtype=np.complex128)  # operator-valued (l, r) matrix
    w[0, 0], w[1, 0], w[2, 0], w[3, 0], w[4, 0] = one, sp, sm, sz, h * sz
    w[4, 1], w[4, 2], w[4, 3], w[4, 4] = 0.5 * J * sm, 0.5 * J * sp, J * sz, one
    bulk = np.transpose(w, (2, 3, 0, 1))
    return [bulk[:, :, 4, :].copy() if k == 0 else (bulk[:, :, :, 0].copy() if k == n - 1 else bulk.copy())
            for k in range(n)]
