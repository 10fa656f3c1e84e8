"""The reference's own hot-path property tests (`/root/reference/test/Ansatz/Chain_test.jl:189-394`)
restated against the oracle, plus dense state-vector known answers for the paths the reference
leaves untested (`Chain_test.jl:396`: evolve!, expect, overlap).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle import chain as oc
from oracle import statevector as sv
from oracle.chain import Chain, MissingSchmidtCoefficientsException, site
from oracle.tenet import Tensor, contract

rng = np.random.default_rng(7)


def rnd(*shape):
    return rng.random(shape) + 1j * rng.random(shape)


def dense(q):
    return q.to_dense()


# Chain_test.jl:189-204
def test_truncate():
    q = Chain([rnd(2, 2), rnd(2, 2, 2), rnd(2, 2)])
    q.canonize_site(site(2), "right", "svd")
    with pytest.raises(MissingSchmidtCoefficientsException):
        q.deepcopy().truncate((site(1), site(2)), maxdim=1)
    t = q.deepcopy().truncate((site(2), site(3)), maxdim=1)
    assert t.tn.size(t.rightindex(site(2))) == 1 and t.tn.size(t.leftindex(site(3))) == 1
    s = q.lambda_between(site(2), site(3)).data
    t = q.deepcopy().truncate((site(2), site(3)), threshold=s[1] + 0.1)
    assert t.tn.size(t.rightindex(site(2))) == 1 and t.tn.size(t.leftindex(site(3))) == 1


# Chain_test.jl:206-221
def test_rand_state():
    q = oc.rand_mps(np.random.default_rng(1), 8, 10)
    assert len(q.outputs()) == 8 and not q.inputs()
    assert np.isclose(q.norm(), 1.0)
    assert max(max(t.shape) for t in q.tn.tensors) <= 10
    # Appendix B: right-canonical on sites 2..n
    for k in range(2, 9):
        assert q.isrightcanonical(site(k))


# Chain_test.jl:241-266
def test_vidal_contract_between():
    q = oc.rand_mps(np.random.default_rng(2), 5, 20)
    can = q.copy().canonize()
    with pytest.raises(ValueError):
        can.copy().contract_between(site(1), site(2), direction="dummy")
    ref = dense(q)
    for i in range(1, 5):
        some = can.copy().contract_between(site(i), site(i + 1))
        assert np.allclose(dense(some), ref)
        assert some.lambda_between(site(i), site(i + 1)) is None
        assert some.isrightcanonical(site(i))
        assert can.copy().contract_between(site(i), site(i + 1), direction="right").isleftcanonical(site(i + 1))
        g = can.tensor_at(site(i))
        lam = can.lambda_between(site(i), site(i + 1))
        b = some.tensor_at(site(i))
        want = contract(g, lam, dims=())
        assert np.allclose(b.permute(want.inds).data, want.data)


# Chain_test.jl:268-306
@pytest.mark.parametrize("method", ["qr", "svd"])
def test_canonize_site(method):
    q = Chain([rnd(4, 4), rnd(4, 4, 4), rnd(4, 4)])
    with pytest.raises(ValueError):
        q.deepcopy().canonize_site(site(1), "left")
    with pytest.raises(ValueError):
        q.deepcopy().canonize_site(site(3), "right")
    ref = dense(q)
    for s, d in [(1, "right"), (2, "right"), (2, "left"), (3, "left")]:
        c = q.deepcopy().canonize_site(site(s), d, method)
        assert c.isleftcanonical(site(s)) if d == "right" else c.isrightcanonical(site(s))
        assert np.allclose(dense(c), ref)
    assert len(q.deepcopy().canonize_site(site(2), "left", "svd").tn) == 4


# Chain_test.jl:308-356
def test_canonize():
    q = Chain([rnd(4, 4), rnd(4, 4, 4), rnd(4, 4, 4), rnd(4, 4, 4), rnd(4, 4)])
    can = q.copy().canonize()
    assert len(can.tn) == 9
    assert np.allclose(dense(can), dense(q))
    assert np.isclose(q.norm(), can.norm())
    lams = can.lambdas()
    assert np.allclose([np.sum(np.abs(l) ** 2) for l in lams], can.norm() ** 2)
    for i in range(1, 6):
        c = q.copy().canonize()
        if i == 1:
            assert c.isleftcanonical(site(1))
        else:
            c.contract_between(site(i - 1), site(i), direction="right")
            if i == 5:
                t = c.tensor_at(site(5))
                c.tn.replace_tensor(t, Tensor(t.data / c.norm(), t.inds))
            assert c.isleftcanonical(site(i))
    for i in range(1, 6):
        c = q.copy().canonize()
        if i == 5:
            assert c.isrightcanonical(site(5))
        else:
            c.contract_between(site(i), site(i + 1), direction="left")
            if i == 1:
                t = c.tensor_at(site(1))
                c.tn.replace_tensor(t, Tensor(t.data / c.norm(), t.inds))
            assert c.isrightcanonical(site(i))


# Chain_test.jl:358-381
def test_mixed_canonize_and_normalize():
    q = Chain([rnd(4, 4), rnd(4, 4, 4), rnd(4, 4, 4), rnd(4, 4, 4), rnd(4, 4)])
    c = q.deepcopy().mixed_canonize(site(3))
    assert len(c.tn) == len(q.tn) + 1
    assert c.isleftcanonical(site(1)) and c.isleftcanonical(site(2))
    assert c.isrightcanonical(site(3)) and c.isrightcanonical(site(4)) and c.isrightcanonical(site(5))
    assert np.allclose(dense(c), dense(q))
    c = q.deepcopy().normalize(site(3))
    assert np.isclose(c.norm(), 1.0)


# Chain_test.jl:383-394
def test_adjoint():
    q = oc.rand_mps(np.random.default_rng(3), 5, 10)
    a = q.adjoint()
    for i in range(1, 6):
        if i < 5:
            assert a.rightindex(site(i, True)) == q.rightindex(site(i)) + "'"
        if i > 1:
            assert a.leftindex(site(i, True)) == q.leftindex(site(i)) + "'"
    assert np.allclose(a.tn.contract().permute([a.sites[site(k, True)] for k in range(1, 6)]).data,
                       np.conj(q.tn.contract().permute([q.sites[site(k)] for k in range(1, 6)]).data))


# ---- beyond the reference's tests (parity unpinned there): dense state-vector known answers ----
def test_overlap_and_expect_dense():
    n = 6
    a = oc.rand_mps(np.random.default_rng(4), n, 8)
    b = oc.rand_mps(np.random.default_rng(5), n, 8)
    va, vb = dense(a), dense(b)
    assert np.isclose(a.overlap(b), np.vdot(vb, va))
    assert np.isclose(a.overlap(a), 1.0)
    Z = np.diag([1.0, -1.0]).astype(complex)
    assert np.isclose(a.expect([oc.gate(Z, [3])]), sv.expect(va, Z, [3], n))
    U = oc.haar_unitary(np.random.default_rng(6))
    assert np.isclose(a.expect([oc.gate(U, [2, 3])]), sv.expect(va, U, [2, 3], n))
    can = a.copy().canonize()
    assert np.isclose(can.expect([oc.gate(Z, [3])]), sv.expect(va, Z, [3], n))
    assert np.isclose(a.expect([oc.gate(np.eye(2), [1])]), a.norm() ** 2)


@pytest.mark.parametrize("iscanonical", [True, False])
def test_evolve_matches_statevector(iscanonical):
    n = 6
    q = oc.rand_mps(np.random.default_rng(8), n, 8)
    psi = dense(q)
    if iscanonical:
        q = q.canonize()
    g = np.random.default_rng(9)
    ntens = len(q.tn)
    for bond in [1, 3, 5, 2, 4]:
        U = oc.haar_unitary(g)
        q.evolve(oc.gate(U, [bond, bond + 1]), iscanonical=iscanonical)
        psi = sv.apply_gate(psi, U, [bond, bond + 1], n)
        if iscanonical:
            assert len(q.tn) == ntens  # Vidal form keeps 2n-1 tensors
    H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    q.evolve(oc.gate(H, [4]))
    psi = sv.apply_gate(psi, H, [4], n)
    assert np.allclose(dense(q), psi, atol=1e-12)
    if iscanonical:
        for lam in q.lambdas():
            assert np.isclose(np.sum(lam ** 2), 1.0)


def test_truncation_weight():
    n = 10
    q = oc.rand_mps(np.random.default_rng(10), n, 8).canonize()
    U = oc.haar_unitary(np.random.default_rng(11))
    full = q.copy().evolve(oc.gate(U, [5, 6]), iscanonical=True)
    s = full.lambdas()[4]
    assert s.shape[0] == 16
    tr = q.copy().evolve(oc.gate(U, [5, 6]), iscanonical=True, maxdim=8)
    kept = tr.lambdas()[4]
    assert kept.shape[0] == 8 and np.allclose(kept, s[:8])
    assert np.isclose(tr.norm() ** 2, np.sum(s[:8] ** 2))
    ren = q.copy().evolve(oc.gate(U, [5, 6]), iscanonical=True, maxdim=8, renormalize=True)
    assert np.isclose(np.sum(ren.lambdas()[4] ** 2), 1.0)
    # identity gate leaves every Λ unchanged
    same = q.copy().evolve(oc.gate(np.eye(4), [5, 6]), iscanonical=True)
    for x, y in zip(same.lambdas(), q.lambdas()):
        assert np.allclose(x[: len(y)], y) and np.allclose(x[len(y):], 0, atol=1e-13)


def test_mpo_composition_against_dense():
    """MPO x MPS has no function in the reference: the oracle's composition vs dense algebra."""
    n = 6
    mpo = oc.heisenberg_mpo_arrays(n)
    H = oc.mpo_to_dense(mpo)
    sz, sp = np.diag([.5, -.5]), np.array([[0, 1], [0, 0]])

    def op(o, k):
        m = np.eye(1)
        for j in range(n):
            m = np.kron(o if j == k else np.eye(2), m)
        return m
    Hd = sum(op(sz, k) @ op(sz, k + 1) + 0.5 * (op(sp, k) @ op(sp.T, k + 1) + op(sp.T, k) @ op(sp, k + 1))
             for k in range(n - 1))
    assert np.allclose(H, Hd)
    arrays = oc.rand_mps_arrays(np.random.default_rng(0), n, 8)
    psi = oc.Chain(arrays)
    v = psi.to_dense()
    assert np.isclose(oc.expect_mpo(psi, mpo), np.vdot(v, Hd @ v))
    assert np.isclose(oc.expect_mpo(psi.copy().canonize(), mpo), np.vdot(v, Hd @ v))
    hp = oc.Chain(oc.apply_mpo_arrays(arrays, mpo))
    assert np.allclose(hp.to_dense(), Hd @ v)
    assert np.allclose(oc.compress(hp.copy()).to_dense(), Hd @ v)
    c = oc.compress(hp.copy(), maxdim=6)
    assert max(len(l) for l in c.lambdas()) == 6


# Chain_test.jl:2-188 (constructors: State / Operator x Open / Periodic, orders, site map, left/right sites)
import _chain_ctor_cases as ctor  # noqa: E402  (tests/ is on sys.path: rootdir-relative import mode)


@pytest.mark.parametrize("case", ctor.ALL, ids=lambda f: f.__name__)
def test_chain_constructors(case):
    case(lambda arrays, **kw: Chain(arrays, **kw), site)


def test_periodic_ring_contractions():
    ctor.check_periodic_ring_contractions(lambda arrays, **kw: Chain(arrays, **kw), site, lambda q: q.to_dense())


# Chain_test.jl:223-235: rand(Chain, Open, Operator; n, p, χ)
@pytest.mark.parametrize("n,chi", [(8, 10), (5, 3), (7, 100), (2, 4)])
def test_rand_operator(n, chi):
    q = oc.rand_mpo(np.random.default_rng(3), n, chi)
    assert q.socket == "operator" and len(q.inputs()) == n and len(q.outputs()) == n
    assert set(q.sites) == {site(i, d) for i in range(1, n + 1) for d in (False, True)} and q.boundary == "open"
    assert np.isclose(q.norm(), 1.0)
    phys = set(q.sites.values())
    assert max(q.tn.size(i) for t in q.tn.tensors for i in t.inds if i not in phys) <= chi
    # default eltype is Float64 (Chain.jl:264) and a complex eltype works too
    assert all(np.isrealobj(t.data) for t in q.tn.tensors)
    assert np.isclose(oc.rand_mpo(np.random.default_rng(4), n, chi, dtype=np.complex128).norm(), 1.0)


# test/Quantum_test.jl:1-45 (sockets, site bookkeeping, error detection) and test/Site_test.jl (sites are (id, dual))
def test_quantum_sockets_and_errors():
    from oracle.chain import Quantum
    from oracle.tenet import TensorNetwork
    q = Quantum(TensorNetwork([Tensor(np.zeros(2), ["i"])]), {site(1): "i"})
    assert (len(q.inputs()), len(q.outputs()), set(q.sites), q.socket_type()) == (0, 1, {site(1)}, "state")
    q = Quantum(TensorNetwork([Tensor(np.zeros(2), ["i"])]), {site(1, True): "i"})
    assert (len(q.inputs()), len(q.outputs()), set(q.sites), q.socket_type()) == (1, 0, {site(1, True)}, "state'")
    q = Quantum(TensorNetwork([Tensor(np.zeros((2, 2)), ["i", "j"])]), {site(1): "i", site(1, True): "j"})
    assert (len(q.inputs()), len(q.outputs()), q.socket_type()) == (1, 1, "operator")
    q = Quantum(TensorNetwork([Tensor(np.zeros(()), [])]), {})
    assert (len(q.inputs()), len(q.outputs()), q.socket_type()) == (0, 0, "scalar") and not q.sites
    tn = TensorNetwork([Tensor(np.zeros(2), ["i"]), Tensor(np.zeros(2), ["i"])])
    with pytest.raises(RuntimeError):
        Quantum(tn, {site(1): "j"})      # index not in the network
    with pytest.raises(RuntimeError):
        Quantum(tn, {site(1): "i"})      # index not open
    s = site(1)
    assert s[0] == 1 and s[1] is False and site(1, True)[1] is True   # Site(id; dual)
    adj = (s[0], not s[1])
    assert adj == site(1, True) and (adj[0], not adj[1]) == s          # adjoint flips dual


# test/Ansatz/Product_test.jl:1-29 + zeros / ones / overlap (Product.jl:48-90)
def test_product_ansatz():
    from oracle.chain import Product
    r = np.random.default_rng(21)
    q = Product([r.random(2) for _ in range(3)])
    assert q.socket_type() == "state" and not q.inputs() and len(q.outputs()) == 3
    assert np.isscalar(q.norm()) or np.ndim(q.norm()) == 0
    assert np.isclose(q.normalize_().norm(), 1.0)
    q = Product([r.random((2, 2)) for _ in range(3)])
    assert q.socket_type() == "operator" and len(q.inputs()) == 3 and len(q.outputs()) == 3
    assert np.ndim(q.norm()) == 0 and np.ndim(q.opnorm()) == 0
    assert np.isclose(q.normalize_().norm(), 1.0)
    z, o = Product.zeros(4), Product.ones(4)
    assert [t.data.tolist() for t in z.tn.tensors] == [[True, False]] * 4
    assert [t.data.tolist() for t in o.tn.tensors] == [[False, True]] * 4
    assert z.overlap(z) == 1 and z.overlap(o) == 0
    a, b = Product([r.random(2) + 1j * r.random(2) for _ in range(3)]), Product([r.random(2) + 1j * r.random(2) for _ in range(3)])
    want = np.prod([np.vdot(x.data, np.conj(y.data)) for x, y in zip(a.tn.tensors, b.tn.tensors)])
    assert np.isclose(a.overlap(b), want)
    with pytest.raises(AssertionError):     # Product.jl:15: no inner indices
        from oracle.chain import Quantum
        from oracle.tenet import TensorNetwork
        Product(_q=Quantum(TensorNetwork([Tensor(np.zeros((2, 2)), ["a", "k"]), Tensor(np.zeros((2, 2)), ["k", "b"])]),
                           {site(1): "a", site(2): "b"}))
    # the product state as a chain (convert(Chain, Product), Chain.jl:174-183) has the same amplitudes
    vs = [r.random(2) + 1j * r.random(2) for _ in range(4)]
    chain = Chain([vs[0].reshape(2, 1)] + [v.reshape(2, 1, 1) for v in vs[1:-1]] + [vs[-1].reshape(2, 1)])
    dense_want = vs[0]
    for v in vs[1:]:
        dense_want = np.kron(v, dense_want)          # site 1 fastest
    assert np.allclose(chain.to_dense(), dense_want)
