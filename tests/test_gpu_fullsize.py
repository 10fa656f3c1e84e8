"""GPU parity at BASELINE.json's full size (chi = 1024, theta 2048 x 2048): one bulk bond of the headline TEBD sweep
checked against LAPACK (the oracle's SVD driver, SciPy gesdd) on the same theta, plus the size-independent
properties of the domain (Vidal form, sum(lambda^2) = norm^2, kept + discarded weight = 1, canonize! leaves the
state unchanged).  n = 22 is the shortest chain whose middle bonds reach 1024 (bond b has dim min(chi, 2^b, 2^(n-b)),
Chain.jl:230-236)."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import chain as oc

pytestmark = pytest.mark.gpu

SIG_TOL = 1e-12
OBS_TOL = 1e-10


@pytest.fixture(scope="module")
def qb():
    import qrochet_b200 as q
    return q


@pytest.fixture(scope="module")
def ctx(qb):
    c = qb.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def state(qb, ctx):
    n, chi = 22, 1024
    arrays = qb.rand_mps_arrays(np.random.default_rng(1000 + 4), n, chi)
    plain = qb.B200MPS(ctx, arrays)
    vidal = plain.copy()
    vidal.canonize()
    return n, chi, plain, vidal


def test_canonize_at_chi_1024_is_vidal_and_keeps_the_state(qb, ctx, state):
    n, chi, plain, vidal = state
    assert vidal.bond_dims() == qb.bond_dims(n, chi)
    assert vidal.bond_dims().count(chi) == 3
    lams = vidal.lambdas()
    for lam in lams:
        assert lam is not None and np.all(np.diff(lam) <= 0) and lam[-1] >= 0
        assert abs(np.sum(lam ** 2) - 1.0) <= OBS_TOL          # sum(lambda^2) = ||psi||^2 on every bond
    # canonize! only changes the gauge
    assert abs(vidal.overlap(plain) - 1.0) <= OBS_TOL
    assert abs(vidal.norm() - 1.0) <= OBS_TOL
    # Lambda_{i-1} Gamma_i left-canonical, Gamma_i Lambda_i right-canonical on the widest site (1024, 2, 1024)
    s = 11
    g = vidal.site(s)
    a = (lams[s - 1][:, None, None] * g).reshape(-1, g.shape[2], order="F")
    assert np.abs(a.conj().T @ a - np.eye(a.shape[1])).max() <= 1e-11
    b = (g * lams[s][None, None, :]).reshape(g.shape[0], -1, order="F")
    assert np.abs(b @ b.conj().T - np.eye(b.shape[0])).max() <= 1e-11


def test_bulk_bond_update_at_chi_1024_matches_lapack(qb, ctx, state):
    n, chi, _, vidal = state
    psi = vidal.copy()
    b = 11                                           # sites (11, 12), 1-based: bonds 10, 11, 12 all have dim 1024
    lams = psi.lambdas()
    gl, gr = psi.site(b - 1), psi.site(b)            # (l, o, r)
    assert gl.shape == (chi, 2, chi) and gr.shape == (chi, 2, chi)
    U = oc.haar_unitary(np.random.default_rng(2000 + b))
    gate = np.reshape(U, (2, 2, 2, 2), order="F")    # (o1, o2, i1, i2)
    # theta = (Lambda_l Gamma_l Lambda)(Gamma_r Lambda_r), gate on (o1, o2)   (Chain.jl:669-685, :635-636)
    left = (lams[b - 2][:, None, None] * gl * lams[b - 1][None, None, :]).reshape(2 * chi, chi, order="F")
    right = (gr * lams[b][None, None, :]).reshape(chi, 2 * chi, order="F")
    theta = (left @ right).reshape(chi, 2, 2, chi, order="F")
    theta = np.einsum("abij,lijr->labr", gate, theta, optimize=True).reshape(2 * chi, 2 * chi, order="F")
    sig = sla.svd(theta, compute_uv=False, lapack_driver="gesdd")
    # truncate! rule (Chain.jl:404-417): i <= min(dim, maxdim) and s_i > threshold (default 1e-16 absolute)
    want_kept = int(np.sum((np.arange(len(sig)) < chi) & (sig > 1e-16)))
    want_dw = float(np.sum(sig[want_kept:] ** 2))

    kept, dw = psi.evolve(gate, [b, b + 1], maxdim=chi, iscanonical=True)
    assert kept == want_kept == chi                   # bit-exact kept count
    got = psi.lambdas()[b - 1]
    assert got.shape == (chi,)
    assert np.abs(got - sig[:chi]).max() <= SIG_TOL * sig[0]
    assert abs(dw - want_dw) <= SIG_TOL
    assert abs(np.sum(got ** 2) + dw - 1.0) <= OBS_TOL            # kept + discarded weight = ||theta||^2 = 1
    assert abs(psi.norm() ** 2 - np.sum(got ** 2)) <= OBS_TOL     # norm^2 after truncate! = kept weight
    # the update keeps the Vidal form (Chain.jl:708-713)
    lams2 = psi.lambdas()
    gl2, gr2 = psi.site(b - 1), psi.site(b)
    a = (lams2[b - 2][:, None, None] * gl2).reshape(-1, chi, order="F")
    assert np.abs(a.conj().T @ a - np.eye(chi)).max() <= 1e-10
    r = (gr2 * lams2[b][None, None, :]).reshape(chi, -1, order="F")
    assert np.abs(r @ r.conj().T - np.eye(chi)).max() <= 1e-10
    # and the truncated theta is the best rank-chi approximation: ||theta - theta_chi||_F^2 = discarded weight
    new = ((lams2[b - 2][:, None, None] * gl2 * got[None, None, :]).reshape(2 * chi, chi, order="F")
           @ (gr2 * lams2[b][None, None, :]).reshape(chi, 2 * chi, order="F"))
    assert abs(np.linalg.norm(theta - new) ** 2 - want_dw) <= 1e-10
