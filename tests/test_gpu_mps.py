"""GPU parity of the fused MPS path (canonize!/mixed_canonize!/truncate!/evolve!/overlap/expect through the
C-ABI) against the CPU oracle on identical seeded inputs.  Comparisons are gauge-invariant (north star):
Schmidt values 1e-12 relative to sigma_1, overlaps / <O> 1e-10 relative, kept counts bit-exact."""
import numpy as np
import pytest

from oracle import chain as oc
from oracle import statevector as sv
from oracle.chain import site

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


def make(qb, ctx, seed, n, chi):
    arrays = oc.rand_mps_arrays(np.random.default_rng(seed), n, chi)
    return oc.Chain(arrays), qb.B200MPS(ctx, arrays)


def dense_from_gpu(g):
    """Contract the device MPS (with its Schmidt vectors) on the host, qubit 1 fastest."""
    lams = g.lambdas()
    n = g.nsites
    psi = np.ones((1, 1), dtype=complex)
    for s in range(n):
        a = g.site(s)  # (l, p, r)
        if s < n - 1 and lams[s] is not None:
            a = a * lams[s][None, None, :]
        psi = np.tensordot(psi, a, axes=(1, 0)).reshape(-1, a.shape[2], order="F")
    return psi[:, 0]


def assert_lams(g_lams, o_lams):
    for b, (x, y) in enumerate(zip(g_lams, o_lams)):
        assert (x is None) == (y is None), b
        if x is None:
            continue
        k = min(len(x), len(y))
        assert np.abs(x[:k] - y[:k]).max() <= SIG_TOL * y[0], b
        assert np.all(np.abs(x[k:]) <= SIG_TOL * y[0]) and np.all(np.abs(y[k:]) <= SIG_TOL * y[0])


def test_roundtrip_and_norm(qb, ctx):
    o, g = make(qb, ctx, 1, 8, 10)
    for a, b in zip(g.arrays(), oc.rand_mps_arrays(np.random.default_rng(1), 8, 10)):
        assert np.array_equal(a, b)
    assert abs(g.norm() - 1.0) < 1e-13  # Chain_test.jl:219
    assert np.allclose(dense_from_gpu(g), o.to_dense(), atol=1e-14)


@pytest.mark.parametrize("n,chi", [(5, 20), (16, 32), (9, 6)])
def test_canonize_matches_oracle(qb, ctx, n, chi):
    o, g = make(qb, ctx, 2, n, chi)
    ref = o.to_dense()
    o = o.canonize()
    g.canonize()
    assert g.form == 1
    assert_lams(g.lambdas(), o.lambdas())
    assert np.allclose(dense_from_gpu(g), ref, atol=1e-12)
    lams = g.lambdas()
    # Chain_test.jl:322-323: sum lambda^2 = |psi|^2 on every bond
    assert np.allclose([np.sum(l ** 2) for l in lams], 1.0, atol=1e-12)
    # Chain_test.jl:325-355: Λ_{i-1}Γ_i left-canonical, Γ_iΛ_i right-canonical
    for s in range(n):
        a = g.site(s)
        al = a if s == 0 else a * lams[s - 1][:, None, None]
        m = al.reshape(-1, a.shape[2], order="F")
        assert np.abs(m.conj().T @ m - np.eye(a.shape[2])).max() < 1e-11
        ar = a if s == n - 1 else a * lams[s][None, None, :]
        m = ar.reshape(a.shape[0], -1, order="F")
        assert np.abs(m @ m.conj().T - np.eye(a.shape[0])).max() < 1e-11


def test_mixed_canonize_and_errors(qb, ctx):
    rng = np.random.default_rng(3)
    arrays = [rng.random((4, 4)) + 0j, rng.random((4, 4, 4)) + 0j, rng.random((4, 4, 4)) + 0j,
              rng.random((4, 4, 4)) + 0j, rng.random((4, 4)) + 0j]
    o = oc.Chain(arrays)
    g = qb.B200MPS(ctx, arrays)
    ref = o.to_dense()
    o.mixed_canonize(site(3))
    g.mixed_canonize(3)
    assert_lams(g.lambdas(), o.lambdas())
    assert [l is not None for l in g.lambdas()] == [False, True, False, False]
    assert np.allclose(dense_from_gpu(g), ref, atol=1e-12)
    for s in range(5):  # Chain_test.jl:364-368
        a = g.site(s)
        if s < 2:
            m = a.reshape(-1, a.shape[2], order="F")
            assert np.abs(m.conj().T @ m - np.eye(a.shape[2])).max() < 1e-12
        else:
            m = a.reshape(a.shape[0], -1, order="F")
            assert np.abs(m @ m.conj().T - np.eye(a.shape[0])).max() < 1e-12
    with pytest.raises(qb.QB200Error):
        qb.B200MPS(ctx, arrays).mixed_canonize(1)  # Site(1) throws in the reference too (Chain.jl:344)


def test_truncate(qb, ctx):
    # Chain_test.jl:189-204
    rng = np.random.default_rng(4)
    arrays = [rng.random((2, 2)) + 0j, rng.random((2, 2, 2)) + 0j, rng.random((2, 2)) + 0j]
    g = qb.B200MPS(ctx, arrays)
    with pytest.raises(qb.MissingSchmidtCoefficientsException):
        g.truncate((1, 2), maxdim=1)
    g.mixed_canonize(3)  # Λ on bond (2,3)
    s = g.lambdas()[1]
    t = g.copy()
    assert t.truncate((2, 3), maxdim=1) == 1 and t.bond_dims()[1] == 1
    t = g.copy()
    assert t.truncate((2, 3), threshold=s[1] + 0.1) == 1 and t.bond_dims()[1] == 1
    t = g.copy()
    assert t.truncate((2, 3), maxdim=5) == 2


@pytest.mark.parametrize("vidal", [True, False])
def test_evolve_matches_oracle_and_statevector(qb, ctx, vidal):
    n = 8
    o, g = make(qb, ctx, 5, n, 8)
    psi = o.to_dense()
    if vidal:
        o.canonize()
        g.canonize()
    rng = np.random.default_rng(6)
    for bond in [1, 3, 5, 7, 2, 4, 6, 4]:
        U = oc.haar_unitary(rng)
        gate = np.reshape(U, (2, 2, 2, 2), order="F")
        o.evolve(oc.gate(U, [bond, bond + 1]), iscanonical=vidal)
        g.evolve(gate, [bond, bond + 1], iscanonical=vidal)
        psi = sv.apply_gate(psi, U, [bond, bond + 1], n)
    H = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    o.evolve(oc.gate(H, [4]))
    g.evolve(H, [4])
    psi = sv.apply_gate(psi, H, [4], n)
    assert np.allclose(dense_from_gpu(g), psi, atol=1e-11)
    assert_lams(g.lambdas(), o.lambdas())
    assert abs(g.norm() - 1.0) < 1e-12


def test_evolve_truncation_counts_and_weights(qb, ctx):
    n = 10
    o, g = make(qb, ctx, 7, n, 8)
    o.canonize()
    g.canonize()
    rng = np.random.default_rng(8)
    for bond in [5, 4, 6, 5]:
        U = oc.haar_unitary(rng)
        full = o.copy().evolve(oc.gate(U, [bond, bond + 1]), iscanonical=True).lambdas()[bond - 1]
        o.evolve(oc.gate(U, [bond, bond + 1]), iscanonical=True, maxdim=8, renormalize=True)
        kept, dw = g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [bond, bond + 1], maxdim=8, iscanonical=True,
                            renormalize=True)
        assert kept == len(o.lambdas()[bond - 1]) == 8            # bit-exact truncation
        assert np.isclose(dw, np.sum(full[8:] ** 2), rtol=1e-9, atol=1e-20)
        assert_lams(g.lambdas(), o.lambdas())
    b = oc.rand_mps(np.random.default_rng(9), n, 4)
    gb = qb.B200MPS(ctx, oc.rand_mps_arrays(np.random.default_rng(9), n, 4))
    want = o.overlap(b)
    got = g.overlap(gb)
    assert abs(got - want) <= OBS_TOL * abs(want)
    # threshold rule
    U = oc.haar_unitary(rng)
    full = o.copy().evolve(oc.gate(U, [3, 4]), iscanonical=True).lambdas()[2]
    thr = float(full[3] * 0.99)
    o.evolve(oc.gate(U, [3, 4]), iscanonical=True, threshold=thr)
    kept, _ = g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [3, 4], threshold=thr, iscanonical=True)
    assert kept == len(o.lambdas()[2]) == int(np.sum(full > thr))


def test_overlap_and_expect(qb, ctx):
    n = 16
    oa, ga = make(qb, ctx, 10, n, 32)
    ob, gb = make(qb, ctx, 11, n, 16)
    want = oa.overlap(ob)
    got = ga.overlap(gb)
    assert abs(got - want) <= OBS_TOL * abs(want)
    assert abs(gb.overlap(ga) - np.conj(want)) <= OBS_TOL * abs(want)
    Z = np.diag([1.0, -1.0]).astype(complex)
    X = np.array([[0, 1], [1, 0]], dtype=complex)
    Y = np.array([[0, -1j], [1j, 0]])
    ops, sites = [Z, X, Y, Z, X], [8, 1, 16, 3, 12]
    want = np.array([oa.expect([oc.gate(o, [s])]) for o, s in zip(ops, sites)])
    got = ga.expect(ops, sites)
    assert np.abs(got - want).max() <= OBS_TOL * max(1.0, np.abs(want).max())
    # config 1 of BASELINE.json: canonize! + overlap + single-site expect, Vidal form keeps <O>
    oa.canonize()
    ga.canonize()
    got2 = ga.expect(ops, sites)
    assert np.abs(got2 - want).max() <= OBS_TOL
    assert abs(ga.overlap(ga) - 1.0) < 1e-12
    assert abs(ga.expect([np.eye(2)], [5])[0] - 1.0) < 1e-12


def test_evolve_layer_equals_sequential_evolve(qb, ctx):
    """The concurrent TEBD layer (independent bond updates on worker streams) is the same as calling evolve!
    bond by bond, and matches the oracle."""
    n = 12
    o, g1 = make(qb, ctx, 21, n, 16)
    o.canonize()
    g1.canonize()
    g2 = g1.copy()
    rng = np.random.default_rng(22)
    for group in ([1, 3, 5, 7, 9, 11], [2, 4, 6, 8, 10], [1, 3, 5, 7, 9, 11]):
        gates = []
        for b in group:
            U = oc.haar_unitary(rng)
            gates.append(np.reshape(U, (2, 2, 2, 2), order="F"))
            o.evolve(oc.gate(U, [b, b + 1]), iscanonical=True, maxdim=16, renormalize=True)
        kept_l, dw_l = g1.evolve_layer(gates, group, maxdim=16, iscanonical=True, renormalize=True)
        seq = [g2.evolve(gt, [b, b + 1], maxdim=16, iscanonical=True, renormalize=True) for gt, b in zip(gates, group)]
        assert kept_l == [k for k, _ in seq]
        assert np.allclose(dw_l, [d for _, d in seq], rtol=1e-9, atol=1e-20)
    for x, y in zip(g1.lambdas(), g2.lambdas()):
        assert np.array_equal(x, y)          # same kernels, same inputs: bit-identical
    assert_lams(g1.lambdas(), o.lambdas())
    with pytest.raises(qb.QB200Error):
        g1.evolve_layer([gates[0], gates[1]], [3, 4], maxdim=16)   # adjacent bonds do not commute


def test_product_state_as_chain_and_its_overlap(qb, ctx):
    """convert(Chain, Product) (Chain.jl:174-183) and overlap(Product, Chain) = <chain|product> (Chain.jl:751-752)
    against dense vectors."""
    from oracle import circuit as ocirc
    n = 9
    rng = np.random.default_rng(41)
    vecs = [rng.standard_normal(2) + 1j * rng.standard_normal(2) for _ in range(n)]
    prod = qb.B200MPS.from_product(ctx, vecs)
    assert prod.bond_dims() == [1] * (n - 1)
    dense_p = ocirc.product_vector(vecs)
    assert np.allclose(dense_from_gpu(prod), dense_p, atol=1e-14)
    o, g = make(qb, ctx, 42, n, 8)
    dense_c = o.to_dense()
    want = np.vdot(dense_c, dense_p)                    # overlap(a, b) = <b|a>
    assert abs(prod.overlap(g) - want) <= OBS_TOL * max(1.0, abs(want))
    assert abs(g.overlap(prod) - np.conj(want)) <= OBS_TOL * max(1.0, abs(want))
    assert abs(prod.norm() - np.linalg.norm(dense_p)) <= OBS_TOL * np.linalg.norm(dense_p)
    # a product state evolves like any chain: one gate, bond 1 -> 2 (rank of the gate across the cut)
    U = oc.haar_unitary(rng)
    prod.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [4, 5])
    assert prod.bond_dims()[3] == 2


def test_evolve_circuit_equals_sequential_evolve(qb, ctx):
    """A gate list in program order with repeated and touching bonds (dependency-scheduled on worker streams) gives
    bit-identical results to the evolve! loop, and matches the oracle."""
    n = 12
    o, g1 = make(qb, ctx, 23, n, 16)
    o.canonize()
    g1.canonize()
    g2 = g1.copy()
    rng = np.random.default_rng(24)
    bonds = list(range(1, n, 2)) + list(range(2, n, 2)) + [5, 5, 6, 4, 1, 11, 10, 2, 3, 7, 8, 9, 1] + list(range(1, n))
    gates = []
    for b in bonds:
        U = oc.haar_unitary(rng)
        gates.append(np.reshape(U, (2, 2, 2, 2), order="F"))
        o.evolve(oc.gate(U, [b, b + 1]), iscanonical=True, maxdim=16, renormalize=True)
    kept_c, dw_c = g1.evolve_circuit(gates, bonds, maxdim=16, iscanonical=True, renormalize=True)
    seq = [g2.evolve(gt, [b, b + 1], maxdim=16, iscanonical=True, renormalize=True) for gt, b in zip(gates, bonds)]
    assert kept_c == [k for k, _ in seq]
    assert np.array_equal(dw_c, [d for _, d in seq])
    for x, y in zip(g1.lambdas(), g2.lambdas()):
        assert np.array_equal(x, y)
    for x, y in zip(g1.arrays(), g2.arrays()):
        assert np.array_equal(x, y)
    assert_lams(g1.lambdas(), o.lambdas())
    assert g1.evolve_circuit([], []) == ([], [])
    with pytest.raises(qb.QB200Error):
        g1.evolve_circuit([gates[0]], [n], maxdim=16)   # bond out of range


def test_mpo_expect_apply_compress_match_oracle(qb, ctx):
    """SURVEY §8 a14 / BASELINE config 3 at test size: <ψ|H|ψ> with the Heisenberg MPO (D = 5), MPO application and
    truncation.  No function exists in the reference for this (parity unpinned there): the composition is defined in
    the oracle and checked against dense linear algebra in tests/test_oracle_properties.py."""
    n, chi = 8, 8
    arrays = oc.rand_mps_arrays(np.random.default_rng(31), n, chi)
    mpo = qb.heisenberg_mpo_arrays(n)
    assert all(np.array_equal(a, b) for a, b in zip(mpo, oc.heisenberg_mpo_arrays(n)))
    o = oc.Chain(arrays)
    g = qb.B200MPS(ctx, arrays)
    want = oc.expect_mpo(o, mpo)
    got = g.expect_mpo(mpo)
    assert abs(got - want) <= OBS_TOL * abs(want)
    # Vidal form and mixed form give the same <H>
    gv = g.copy().canonize()
    assert abs(gv.expect_mpo(mpo) - want) <= OBS_TOL * abs(want)
    gm = g.copy().mixed_canonize(4)
    assert abs(gm.expect_mpo(mpo) - want) <= OBS_TOL * abs(want)
    # H|ψ>: bonds fuse to chi*D, dense vector equals the oracle's
    hg = g.copy().apply_mpo(mpo)
    ho = oc.Chain(oc.apply_mpo_arrays(arrays, mpo))
    assert hg.bond_dims() == [min(2 ** (b + 1), 2 ** (n - b - 1), chi) * 5 for b in range(n - 1)]
    assert np.allclose(dense_from_gpu(hg), ho.to_dense(), atol=1e-12)
    # compress without truncation: same state, Vidal form, Schmidt values as the oracle's
    cg = hg.copy().compress()
    co = oc.compress(ho.copy())
    assert np.allclose(dense_from_gpu(cg), ho.to_dense(), atol=1e-11)
    assert_lams([l[: len(m)] for l, m in zip(cg.lambdas(), co.lambdas())], co.lambdas())
    # compress with truncation: bit-exact kept counts, Schmidt values to 1e-12
    cg = hg.copy().compress(maxdim=6)
    co = oc.compress(ho.copy(), maxdim=6)
    assert cg.bond_dims() == [len(l) for l in co.lambdas()]
    assert_lams(cg.lambdas(), co.lambdas())
    ov = cg.overlap(g)
    assert abs(ov - co.overlap(o)) <= OBS_TOL * abs(ov)


def test_product_state_and_rank_deficient_theta(qb, ctx):
    """Edge case: bond dimension 1 everywhere (|0...0>, `zeros(Product, n)` converted to a Chain, Chain.jl:174-183):
    theta has rank 1 before the gate and rank <= 4 after it.  With an explicit threshold the kept count is
    well-defined and must match the oracle bit-exactly (the reference's default absolute 1e-16 threshold is
    numerically fragile for exactly rank-deficient theta, SURVEY.md §7.3(2), so it is not asserted here)."""
    n = 6
    arrays = [np.array([[1.0], [0.0]], dtype=complex)] + [np.array([[[1.0]], [[0.0]]], dtype=complex)] * (n - 2) + \
             [np.array([[1.0], [0.0]], dtype=complex)]
    o = oc.Chain(arrays).canonize()
    g = qb.B200MPS(ctx, arrays).canonize()
    assert g.bond_dims() == [1] * (n - 1)
    rng = np.random.default_rng(41)
    psi = sv.zero_state(n)
    for bond in [1, 3, 5, 2, 4, 3]:
        U = oc.haar_unitary(rng)
        o.evolve(oc.gate(U, [bond, bond + 1]), iscanonical=True, threshold=1e-12)
        kept, _ = g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [bond, bond + 1], threshold=1e-12, iscanonical=True)
        assert kept == len(o.lambdas()[bond - 1])
        psi = sv.apply_gate(psi, U, [bond, bond + 1], n)
    assert np.allclose(dense_from_gpu(g), psi, atol=1e-12)
    assert_lams(g.lambdas(), o.lambdas())


def test_config2_shape_properties(qb, ctx):
    """BASELINE config 2 at full width (n = 64, chi = 256): one brickwork layer of Haar two-site gates with
    truncation to 256.  The oracle needs minutes at this size, so the checks are the size-independent properties
    of the domain: bit-exact kept counts (= maxdim on bulk bonds), sum lambda^2 = 1 after renormalize, discarded
    weight = 1 - kept weight, isometry of the new Gamma*Lambda pairs, and layer == bond-by-bond evolve."""
    n, chi = 64, 256
    arrays = qb.rand_mps_arrays(np.random.default_rng(1002), n, chi)
    g = qb.B200MPS(ctx, arrays).canonize()
    assert abs(g.norm() - 1.0) < 1e-10
    dims = qb.bond_dims(n, chi)
    assert g.bond_dims() == dims
    odd = list(range(1, n, 2))
    gates = [qb.haar_gate(np.random.default_rng(2000 + b)) for b in odd]
    ref = g.copy()
    kept, dw = g.evolve_layer(gates, odd, maxdim=chi, iscanonical=True, renormalize=True)
    d = [1] + dims + [1]
    assert kept == [min(chi, 2 * d[b - 1], 2 * d[b + 1]) for b in odd]      # bit-exact truncation
    lams = g.lambdas()
    for b, w in zip(odd, dw):
        lam = lams[b - 1]
        assert abs(np.sum(lam ** 2) - 1.0) < 1e-12 and np.all(np.diff(lam) <= 0) and 0.0 <= w < 1.0
    # Gamma_l scaled by its left Lambda is an isometry (left-canonical), Gamma_r by its right Lambda right-canonical
    for b in (1, 31, 63):
        a = g.site(b - 1)
        al = a if b == 1 else a * lams[b - 2][:, None, None]
        m = al.reshape(-1, a.shape[2], order="F")
        assert np.abs(m.conj().T @ m - np.eye(m.shape[1])).max() < 1e-10
        c = g.site(b)
        cr = c if b + 1 == n else c * lams[b][None, None, :]
        m = cr.reshape(c.shape[0], -1, order="F")
        assert np.abs(m @ m.conj().T - np.eye(m.shape[0])).max() < 1e-10
    # same as three bond-by-bond calls
    for b, gt, k in list(zip(odd, gates, kept))[14:17]:
        kk, _ = ref.evolve(gt, [b, b + 1], maxdim=chi, iscanonical=True, renormalize=True)
        assert kk == k and np.array_equal(ref.lambdas()[b - 1], lams[b - 1])
