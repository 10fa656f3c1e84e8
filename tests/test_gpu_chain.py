"""The reference's own hot-path tests (`/root/reference/test/Ansatz/Chain_test.jl:189-394`) run on the DEVICE through
the label-driven mirror of Chain.jl (qrochet_b200.chain): the path Julia takes after `adapt(B200Array, ψ)`, where
every `contract` / `svd!` / `qr!` / `slice!` / `conj` dispatches to the Tenet-level C-ABI entry points.  Plus
parity with the oracle on identical inputs (gauge-invariant quantities)."""
import numpy as np
import pytest

from oracle import chain as oc
from oracle import statevector as sv

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    import qrochet_b200 as q
    return q


@pytest.fixture(scope="module")
def ctx(qb):
    c = qb.Context(0)
    yield c
    c.close()


rng = np.random.default_rng(7)


def rnd(*shape):
    return rng.random(shape) + 1j * rng.random(shape)


def test_truncate(qb, ctx):  # Chain_test.jl:189-204
    ch, site = qb.chain, qb.chain.site
    arrays = [rnd(2, 2), rnd(2, 2, 2), rnd(2, 2)]
    q = ch.Chain(ctx, arrays)
    q.canonize_site(site(2), "right", "svd")
    with pytest.raises(qb.MissingSchmidtCoefficientsException):
        q.copy().truncate((site(1), site(2)), maxdim=1)
    t = q.copy().truncate((site(2), site(3)), maxdim=1)
    assert t.tn.size(t.rightindex(site(2))) == 1 and t.tn.size(t.leftindex(site(3))) == 1
    s = q.lambda_between(site(2), site(3)).to_host()
    t = q.copy().truncate((site(2), site(3)), threshold=s[1] + 0.1)
    assert t.tn.size(t.rightindex(site(2))) == 1


@pytest.mark.parametrize("method", ["qr", "svd"])
def test_canonize_site(qb, ctx, method):  # Chain_test.jl:268-306
    ch, site = qb.chain, qb.chain.site
    arrays = [rnd(4, 4), rnd(4, 4, 4), rnd(4, 4)]
    q = ch.Chain(ctx, arrays)
    with pytest.raises(ValueError):
        q.copy().canonize_site(site(1), "left")
    with pytest.raises(ValueError):
        q.copy().canonize_site(site(3), "right")
    ref = oc.Chain(arrays).to_dense()
    for s, d in [(1, "right"), (2, "right"), (2, "left"), (3, "left")]:
        c = q.copy().canonize_site(site(s), d, method)
        assert c.isleftcanonical(site(s)) if d == "right" else c.isrightcanonical(site(s))
        assert np.allclose(c.to_dense(), ref, atol=1e-12)
    assert len(q.copy().canonize_site(site(2), "left", "svd").tn) == 4


def test_canonize_mixed_normalize(qb, ctx):  # Chain_test.jl:308-381
    ch, site = qb.chain, qb.chain.site
    arrays = [rnd(4, 4), rnd(4, 4, 4), rnd(4, 4, 4), rnd(4, 4, 4), rnd(4, 4)]
    q = ch.Chain(ctx, arrays)
    o = oc.Chain(arrays)
    can = q.copy().canonize()
    ocan = o.copy().canonize()
    assert len(can.tn) == 9
    assert np.allclose(can.to_dense(), o.to_dense(), atol=1e-12)
    assert np.isclose(q.norm(), can.norm())
    for x, y in zip(can.lambdas(), ocan.lambdas()):
        assert np.abs(x - y).max() <= 1e-12 * y[0]
    assert np.allclose([np.sum(l ** 2) for l in can.lambdas()], can.norm() ** 2)
    for i in range(2, 5):
        c = q.copy().canonize()
        c.contract_between(site(i - 1), site(i), direction="right")
        assert c.isleftcanonical(site(i))
        c = q.copy().canonize()
        c.contract_between(site(i), site(i + 1), direction="left")
        assert c.isrightcanonical(site(i))
    m = q.copy().mixed_canonize(site(3))
    assert len(m.tn) == len(q.tn) + 1
    assert m.isleftcanonical(site(1)) and m.isleftcanonical(site(2))
    assert m.isrightcanonical(site(3)) and m.isrightcanonical(site(4)) and m.isrightcanonical(site(5))
    assert np.allclose(m.to_dense(), o.to_dense(), atol=1e-12)
    assert np.isclose(q.copy().normalize(site(3)).norm(), 1.0)


def test_adjoint_overlap_expect(qb, ctx):  # Chain_test.jl:383-394 + the untested evolve!/expect/overlap (:396)
    ch, site = qb.chain, qb.chain.site
    n = 6
    a_arr = oc.rand_mps_arrays(np.random.default_rng(4), n, 8)
    b_arr = oc.rand_mps_arrays(np.random.default_rng(5), n, 8)
    a, b = ch.Chain(ctx, a_arr), ch.Chain(ctx, b_arr)
    oa, ob = oc.Chain(a_arr), oc.Chain(b_arr)
    adj = a.adjoint()
    for i in range(1, n):
        assert adj.rightindex(site(i, True)) == a.rightindex(site(i)) + "'"
    assert abs(a.overlap(b) - oa.overlap(ob)) < 1e-12
    assert abs(a.norm() - 1.0) < 1e-12
    Z = np.diag([1.0, -1.0]).astype(complex)
    assert abs(a.expect([(Z, [3])]) - oa.expect([oc.gate(Z, [3])])) < 1e-12
    U = oc.haar_unitary(np.random.default_rng(6))
    G = np.reshape(U, (2, 2, 2, 2), order="F")
    assert abs(a.expect([(G, [2, 3])]) - sv.expect(oa.to_dense(), U, [2, 3], n)) < 1e-12
    assert abs(a.copy().canonize().expect([(Z, [3])]) - oa.expect([oc.gate(Z, [3])])) < 1e-11


@pytest.mark.parametrize("iscanonical", [True, False])
def test_evolve_generic_path_matches_statevector_and_fused_path(qb, ctx, iscanonical):
    ch = qb.chain
    n = 6
    arrays = oc.rand_mps_arrays(np.random.default_rng(8), n, 8)
    q = ch.Chain(ctx, arrays)
    fused = qb.B200MPS(ctx, arrays)
    psi = oc.Chain(arrays).to_dense()
    if iscanonical:
        q.canonize()
        fused.canonize()
    g = np.random.default_rng(9)
    ntens = len(q.tn)
    for bond in [1, 3, 5, 2, 4]:
        U = oc.haar_unitary(g)
        G = np.reshape(U, (2, 2, 2, 2), order="F")
        q.evolve(G, [bond, bond + 1], iscanonical=iscanonical, maxdim=8)
        fused.evolve(G, [bond, bond + 1], iscanonical=iscanonical, maxdim=8)
        psi = sv.apply_gate(psi, U, [bond, bond + 1], n)
        if iscanonical:
            assert len(q.tn) == ntens
    assert np.allclose(q.to_dense(), psi, atol=1e-11)        # chi = 8 holds a 6-qubit state exactly
    for x, y in zip(q.lambdas(), fused.lambdas()):            # generic (Tenet-level) and fused paths agree
        assert (x is None) == (y is None)
        if x is not None:
            k = min(len(x), len(y))
            assert np.abs(x[:k] - y[:k]).max() <= 1e-12 * max(y[0], 1e-300)


# ---- Chain_test.jl:2-188 on device arrays: State / Operator x Open / Periodic constructors, site map, left/right sites
import _chain_ctor_cases as ctor  # noqa: E402


@pytest.mark.parametrize("case", ctor.ALL, ids=lambda f: f.__name__)
def test_chain_constructors_on_device(qb, ctx, case):
    case(lambda arrays, **kw: qb.chain.Chain(ctx, arrays, **kw), qb.chain.site)


def test_periodic_ring_contractions_on_device(qb, ctx):
    ctor.check_periodic_ring_contractions(lambda arrays, **kw: qb.chain.Chain(ctx, arrays, **kw), qb.chain.site,
                                          lambda q: q.to_dense())
