"""The reference's constructor / site-map tests (`/root/reference/test/Ansatz/Chain_test.jl:2-188`) restated once and
run against BOTH implementations of the Chain bookkeeping: the oracle (`oracle.chain.Chain`, CPU) and the label-driven
device mirror (`qrochet_b200.chain.Chain`, GPU).  `make(arrays, **kw)` builds a chain; `site` is the Site ctor."""
import numpy as np

rng = np.random.default_rng(11)


def rnd(*shape):
    return rng.random(shape) + 1j * rng.random(shape)


def shape_at(q, s):
    return tuple(q.tensor_at(s).shape)


def check_periodic_state(make, site):  # Chain_test.jl:3-37
    q = make([rnd(2, 4, 4) for _ in range(3)], boundary="periodic")
    assert q.socket == "state" and not q.inputs() and len(q.outputs()) == 3
    assert set(q.sites) == {site(1), site(2), site(3)} and q.boundary == "periodic"
    assert q.leftindex(site(1)) == q.rightindex(site(3)) is not None
    arrays = [rnd(2, 1, 4), rnd(2, 4, 3), rnd(2, 3, 1)]
    q = make(arrays, boundary="periodic")  # default order (o, l, r)
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(2, 1, 4), (2, 4, 3), (2, 3, 1)]
    for a, b in ((1, 3), (2, 1), (3, 2)):
        assert q.leftindex(site(a)) == q.rightindex(site(b))
    q = make([np.transpose(a, (2, 0, 1)) for a in arrays], boundary="periodic", order=("r", "o", "l"))
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(4, 2, 1), (3, 2, 4), (1, 2, 3)]
    for a, b in ((1, 3), (2, 1), (3, 2)):
        assert q.leftindex(site(a)) == q.rightindex(site(b))
    for i in (1, 2, 3):
        assert q.tn.size(q.sites[site(i)]) == 2


def check_periodic_operator(make, site):  # Chain_test.jl:39-79
    q = make([rnd(2, 2, 4, 4) for _ in range(3)], boundary="periodic", socket="operator")
    assert q.socket == "operator" and len(q.inputs()) == 3 and len(q.outputs()) == 3
    assert set(q.sites) == {site(i, d) for i in (1, 2, 3) for d in (False, True)} and q.boundary == "periodic"
    assert q.leftindex(site(1)) == q.rightindex(site(3)) is not None
    arrays = [rnd(2, 4, 1, 3), rnd(2, 4, 3, 6), rnd(2, 4, 6, 1)]  # default order (o, i, l, r)
    q = make(arrays, boundary="periodic", socket="operator")
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(2, 4, 1, 3), (2, 4, 3, 6), (2, 4, 6, 1)]
    for a, b in ((1, 3), (2, 1), (3, 2)):
        assert q.leftindex(site(a)) == q.rightindex(site(b))
    for i in (1, 2, 3):
        assert q.tn.size(q.sites[site(i)]) == 2 and q.tn.size(q.sites[site(i, True)]) == 4
    q = make([np.transpose(a, (3, 0, 2, 1)) for a in arrays], boundary="periodic", socket="operator",
             order=("r", "o", "l", "i"))
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(3, 2, 1, 4), (6, 2, 3, 4), (1, 2, 6, 4)]
    for a, b in ((1, 3), (2, 1), (3, 2)):
        assert q.leftindex(site(a)) == q.rightindex(site(b)) is not None
    for i in (1, 2, 3):
        assert q.tn.size(q.sites[site(i)]) == 2 and q.tn.size(q.sites[site(i, True)]) == 4


def check_open_state(make, site):  # Chain_test.jl:83-116
    q = make([rnd(2, 2), rnd(2, 2, 2), rnd(2, 2)])
    assert q.socket == "state" and not q.inputs() and len(q.outputs()) == 3 and q.boundary == "open"
    assert q.leftindex(site(1)) is None and q.rightindex(site(3)) is None
    arrays = [rnd(2, 1), rnd(2, 1, 3), rnd(2, 3)]
    q = make(arrays)
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(2, 1), (2, 1, 3), (2, 3)]
    assert q.leftindex(site(1)) is None and q.rightindex(site(3)) is None
    assert q.leftindex(site(2)) == q.rightindex(site(1)) and q.leftindex(site(3)) == q.rightindex(site(2))
    q = make([arrays[0].T, np.transpose(arrays[1], (2, 0, 1)), arrays[2]], order=("r", "o", "l"))
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(1, 2), (3, 2, 1), (2, 3)]
    assert q.leftindex(site(1)) is None and q.rightindex(site(3)) is None
    assert q.leftindex(site(2)) == q.rightindex(site(1)) is not None
    assert q.leftindex(site(3)) == q.rightindex(site(2)) is not None
    for i in (1, 2, 3):
        assert q.tn.size(q.sites[site(i)]) == 2


def check_open_operator(make, site):  # Chain_test.jl:117-160
    q = make([rnd(2, 2, 4), rnd(2, 2, 4, 4), rnd(2, 2, 4)], socket="operator")
    assert q.socket == "operator" and len(q.inputs()) == 3 and len(q.outputs()) == 3 and q.boundary == "open"
    assert q.leftindex(site(1)) is None and q.rightindex(site(3)) is None
    arrays = [rnd(2, 4, 1), rnd(2, 4, 1, 3), rnd(2, 4, 3)]  # default order (o, i, l, r)
    q = make(arrays, socket="operator")
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(2, 4, 1), (2, 4, 1, 3), (2, 4, 3)]
    assert q.leftindex(site(2)) == q.rightindex(site(1)) is not None
    assert q.leftindex(site(3)) == q.rightindex(site(2)) is not None
    for i in (1, 2, 3):
        assert q.tn.size(q.sites[site(i)]) == 2 and q.tn.size(q.sites[site(i, True)]) == 4
    q = make([np.transpose(arrays[0], (2, 0, 1)), np.transpose(arrays[1], (3, 0, 2, 1)), np.transpose(arrays[2], (0, 2, 1))],
             socket="operator", order=("r", "o", "l", "i"))
    assert [shape_at(q, site(i)) for i in (1, 2, 3)] == [(1, 2, 4), (3, 2, 1, 4), (2, 3, 4)]
    assert q.leftindex(site(1)) is None and q.rightindex(site(3)) is None
    assert q.leftindex(site(2)) == q.rightindex(site(1)) is not None


def check_sites(make, site):  # Chain_test.jl:165-188
    q = make([rnd(2, 4, 4) for _ in range(3)], boundary="periodic")
    assert [q.leftsite(site(i)) for i in (1, 2, 3)] == [site(3), site(1), site(2)]
    assert [q.rightsite(site(i)) for i in (1, 2, 3)] == [site(2), site(3), site(1)]
    q = make([rnd(2, 2), rnd(2, 2, 2), rnd(2, 2)])
    assert q.leftsite(site(1)) is None and q.rightsite(site(3)) is None
    assert q.leftsite(site(2)) == site(1) and q.leftsite(site(3)) == site(2)
    assert q.rightsite(site(2)) == site(3) and q.rightsite(site(1)) == site(2)


def check_bad_arguments(make, site):  # the assertions / ArgumentErrors of Chain.jl:37-40,65-69
    import pytest
    with pytest.raises((AssertionError, ValueError)):
        make([rnd(2, 4), rnd(2, 4, 4), rnd(2, 4, 4)], boundary="periodic")       # all arrays need 3 dims
    with pytest.raises((AssertionError, ValueError)):
        make([rnd(2, 2, 2), rnd(2, 2, 2), rnd(2, 2)])                             # open: first array 2 dims
    with pytest.raises(ValueError):
        make([rnd(2, 2), rnd(2, 2, 2), rnd(2, 2)], order=("o", "l", "x"))


def check_periodic_ring_contractions(make, site, dense_of):
    """A periodic MPS is a ring: norm / overlap / adjoint go through the generic network contraction; compared with
    the dense state  psi[s1..sn] = tr(A1[s1] ... An[sn])."""
    n, chi = 5, 3
    a = [rnd(2, chi, chi) for _ in range(n)]
    b = [rnd(2, chi, chi) for _ in range(n)]

    def ring_state(arrs):
        psi = np.zeros((2,) * n, complex)
        for idx in np.ndindex(*(2,) * n):
            m = np.eye(chi, dtype=complex)
            for k in range(n):
                m = m @ arrs[k][idx[k]]
            psi[idx] = np.trace(m)
        return np.reshape(psi, -1, order="F")

    pa, pb = ring_state(a), ring_state(b)
    qa, qb_ = make(a, boundary="periodic"), make(b, boundary="periodic")
    assert np.allclose(dense_of(qa), pa, atol=1e-10 * np.linalg.norm(pa))
    assert abs(qa.norm() - np.linalg.norm(pa)) <= 1e-10 * np.linalg.norm(pa)
    assert abs(qa.overlap(qb_) - np.vdot(pb, pa)) <= 1e-10 * abs(np.vdot(pb, pa))
    adj = qa.adjoint()
    assert adj.boundary == "periodic" and set(adj.sites) == {site(i, True) for i in range(1, n + 1)}


ALL = [check_periodic_state, check_periodic_operator, check_open_state, check_open_operator, check_sites,
       check_bad_arguments]
