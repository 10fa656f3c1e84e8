"""GPU parity of the keyword branches of `evolve!` / `expect` the fused path did not cover in round 1
(VERDICT r1 "missing" 6): `iscanonical = false` on a canonized chain (Chain.jl:615,645), `renormalize`
without `iscanonical` (`normalize!` = `mixed_canonize!` + Λ normalisation, Chain.jl:655-656,532-536), physical
dimension != 2 (`p` keyword of `rand`, Chain.jl:226), observables on more than one lane (Chain.jl:724-735), and
`canonize!` of a chain that already carries Schmidt vectors (Chain.jl:372).  Everything is compared with the CPU
oracle on the same seeded inputs; tolerances are the north star's (sigma 1e-12 sigma_1, <O> 1e-10)."""
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


def assert_lams(g_lams, o_lams):
    for b, (x, y) in enumerate(zip(g_lams, o_lams)):
        assert (x is None) == (y is None), b
        if x is None:
            continue
        assert len(x) == len(y), b                                  # kept counts: bit-exact
        assert np.abs(x - y).max() <= SIG_TOL * y[0], b


def gpu_dense(g):
    lams, n = g.lambdas(), g.nsites
    psi = np.ones((1, 1), dtype=complex)
    for s in range(n):
        a = g.site(s)
        if s < n - 1 and lams[s] is not None:
            a = a * lams[s][None, None, :]
        psi = np.tensordot(psi, a, axes=(1, 0)).reshape(-1, a.shape[2], order="F")
    return psi[:, 0]


def test_noncanonical_evolve_on_a_canonized_chain(qb, ctx):
    """The reference's DEFAULT call `evolve!(ψ, G; maxdim)` on a Vidal chain contracts only Γ Λ Γ (no outer Λ): the
    truncation spectrum is NOT the Schmidt spectrum; kept counts and spectra must follow the oracle, not the Vidal rule."""
    n = 8
    arrays = oc.rand_mps_arrays(np.random.default_rng(31), n, 16)
    o = oc.Chain(arrays).canonize()
    g = qb.B200MPS(ctx, arrays).canonize()
    ref = o.to_dense()
    rng = np.random.default_rng(32)
    for bond in (4, 2, 5, 4):
        U = oc.haar_unitary(rng)
        o.evolve(oc.gate(U, [bond, bond + 1]), maxdim=6)
        kept, _ = g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [bond, bond + 1], maxdim=6)
        assert kept == len(o.lambdas()[bond - 1])
        assert_lams(g.lambdas(), o.lambdas())
    assert g.form == 0
    assert np.allclose(gpu_dense(g), o.to_dense(), atol=1e-11)
    # a later Vidal-style call still follows the oracle (it uses whatever Schmidt vectors sit on the bonds)
    U = oc.haar_unitary(rng)
    o.evolve(oc.gate(U, [6, 7]), iscanonical=True, maxdim=8)
    g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [6, 7], iscanonical=True, maxdim=8)
    assert_lams(g.lambdas(), o.lambdas())
    assert np.allclose(gpu_dense(g), o.to_dense(), atol=1e-11)
    # canonize! of a chain that carries Schmidt vectors absorbs them (contract!(tn, virtualind), Chain.jl:372)
    oc_ = o.copy().canonize()
    g.canonize()
    assert g.form == 1
    assert_lams(g.lambdas(), oc_.lambdas())
    assert np.allclose(gpu_dense(g), oc_.to_dense(), atol=1e-11)
    assert abs(np.vdot(ref, ref) - 1.0) < 1e-12


def test_noncanonical_renormalize_is_normalize(qb, ctx):
    """`renormalize && !iscanonical` -> `normalize!(ψ, bond[1])` (Chain.jl:655-656)."""
    n = 7
    arrays = oc.rand_mps_arrays(np.random.default_rng(41), n, 8)
    o, g = oc.Chain(arrays), qb.B200MPS(ctx, arrays)
    rng = np.random.default_rng(42)
    for bond in (3, 5):
        U = oc.haar_unitary(rng)
        o.evolve(oc.gate(U, [bond, bond + 1]), maxdim=4, renormalize=True)
        kept, _ = g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [bond, bond + 1], maxdim=4, renormalize=True)
        assert kept == 4
        assert_lams(g.lambdas(), o.lambdas())
        assert abs(g.norm() - 1.0) <= OBS_TOL and abs(o.norm() - 1.0) <= OBS_TOL
        want = o.to_dense()
        got = gpu_dense(g)
        assert abs(abs(np.vdot(want, got)) - 1.0) <= OBS_TOL          # same state up to a global phase
    # gate lists take the same route (sequential: normalize! re-canonizes the whole chain)
    gates = [np.reshape(oc.haar_unitary(rng), (2, 2, 2, 2), order="F") for _ in range(2)]
    for gt, b in zip(gates, (2, 4)):
        o.evolve(oc.Dense(gt, [site(b), site(b + 1), site(b, True), site(b + 1, True)]), maxdim=4, renormalize=True)
    g.evolve_circuit(gates, [2, 4], maxdim=4, renormalize=True)
    assert_lams(g.lambdas(), o.lambdas())
    # on the first bond the reference throws (mixed_canonize!(ψ, Site(1)), Chain.jl:344)
    with pytest.raises(ValueError):
        o.copy().evolve(oc.gate(np.eye(4), [1, 2]), maxdim=4, renormalize=True)
    with pytest.raises(qb.QB200Error):
        g.evolve(np.reshape(np.eye(4), (2, 2, 2, 2)), [1, 2], maxdim=4, renormalize=True)


@pytest.mark.parametrize("vidal", [False, True])
def test_expect_with_multi_lane_observables(qb, ctx, vidal):
    """`expect(ψ, observables)` with a mixed list of 1- and 2-lane observables (Chain.jl:724-735)."""
    n = 9
    arrays = oc.rand_mps_arrays(np.random.default_rng(51), n, 12)
    o, g = oc.Chain(arrays), qb.B200MPS(ctx, arrays)
    if vidal:
        o.canonize()
        g.canonize()
    rng = np.random.default_rng(52)
    Z = np.diag([1.0, -1.0]).astype(complex)
    A = rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4))
    B = oc.haar_unitary(rng)
    obs_o = [oc.gate(A, [3, 4]), oc.gate(Z, [7]), oc.gate(B, [6, 5])]
    obs_g = [(np.reshape(A, (2, 2, 2, 2), order="F"), [3, 4]), (Z, [7]), (np.reshape(B, (2, 2, 2, 2), order="F"), [6, 5])]
    want = o.expect(obs_o)
    got = g.expect_observables(obs_g)
    assert abs(got - want) <= OBS_TOL * max(1.0, abs(want))
    # dense cross-check: <psi| B_{6,5} Z_7 A_{3,4} |psi>
    psi = oc.Chain(arrays).to_dense()
    phi = sv.apply_gate(psi, A, [3, 4], n)
    phi = sv.apply_gate(phi, Z, [7], n)
    phi = sv.apply_gate(phi, B, [6, 5], n)
    assert abs(got - np.vdot(psi, phi)) <= OBS_TOL * max(1.0, abs(want))
    assert g.form == (1 if vidal else 0)                                # ψ itself is untouched


def test_physical_dimension_three(qb, ctx):
    """`rand(Chain, Open, State; n, χ, p = 3)` (Chain.jl:226) through canonize! / evolve! / overlap / expect."""
    n, chi, p = 6, 9, 3
    arrays = oc.rand_mps_arrays(np.random.default_rng(61), n, chi, p=p)
    o, g = oc.Chain(arrays), qb.B200MPS(ctx, arrays)
    assert abs(g.norm() - 1.0) <= OBS_TOL
    o.canonize()
    g.canonize()
    assert_lams(g.lambdas(), o.lambdas())
    rng = np.random.default_rng(62)
    for bond in (3, 2, 4):
        U = oc.haar_unitary(rng, d=p * p)
        G = np.reshape(U, (p, p, p, p), order="F")
        o.evolve(oc.gate(U, [bond, bond + 1]), iscanonical=True, maxdim=chi, renormalize=True)
        kept, _ = g.evolve(G, [bond, bond + 1], iscanonical=True, maxdim=chi, renormalize=True)
        assert kept == len(o.lambdas()[bond - 1])
    assert_lams(g.lambdas(), o.lambdas())
    S = np.diag([1.0, 0.0, -1.0]).astype(complex)
    want = o.expect([oc.gate(S, [4])])
    assert abs(g.expect([S], [4])[0] - want) <= OBS_TOL
    assert abs(g.expect_observables([(S, [4])]) - want) <= OBS_TOL
    other = oc.rand_mps_arrays(np.random.default_rng(63), n, 4, p=p)
    want = o.overlap(oc.Chain(other))
    assert abs(g.overlap(qb.B200MPS(ctx, other)) - want) <= OBS_TOL


def test_oracle_restarts_from_device_state(qb, ctx):
    """`oracle.chain.chain_from_vidal` (the hand-over the full-size tests use): the oracle built from the device's
    Vidal state is the same state and evolves to the same Schmidt values."""
    n = 8
    arrays = oc.rand_mps_arrays(np.random.default_rng(71), n, 16)
    g = qb.B200MPS(ctx, arrays).canonize()
    o = oc.chain_from_vidal([g.site(s) for s in range(n)], g.lambdas())
    assert np.allclose(o.to_dense(), gpu_dense(g), atol=1e-13)
    U = oc.haar_unitary(np.random.default_rng(72))
    o.evolve(oc.gate(U, [4, 5]), iscanonical=True, maxdim=8, renormalize=True)
    g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [4, 5], iscanonical=True, maxdim=8, renormalize=True)
    assert_lams(g.lambdas(), o.lambdas())


def test_circuit_front_end_drives_the_chain(qb, ctx):
    """`evolve!(::Ansatz, ::Gate)` (ext/QrochetQuacExt.jl:15) and a gate list read from the text format: a nearest-
    neighbour circuit on |0..0> as a chain, against the state vector (threshold 1e-14 keeps the bonds at their rank)."""
    from qrochet_b200 import gates as G
    n = 6
    text = "qubits 6\nh 1\ncx 1 2\ncx 2 3\nrz 3 0.4\nfsim 3 4 0.3 0.9\nh 6\ncx 6 5\nrzz 4 5 0.7\nswap 2 3\nu3 2 0.1 0.2 0.3\ncz 2 1\n"
    nq, circ = G.loads(text)
    assert nq == n
    e0 = np.array([1.0, 0.0])
    psi = qb.B200MPS.from_product(ctx, [e0] * n)
    ref = sv.zero_state(n)
    for g in circ:
        G.evolve_gate(psi, g, threshold=1e-14)
        ref = sv.apply_gate(ref, g.matrix(), list(g.lanes), n)
    assert np.allclose(gpu_dense(psi), ref, atol=1e-12)
    # the same circuit as ONE dependency-scheduled gate list per run of two-lane gates
    psi2 = qb.B200MPS.from_product(ctx, [e0] * n)
    G.evolve_gates(psi2, circ, threshold=1e-14)
    assert np.allclose(gpu_dense(psi2), ref, atol=1e-12)
    assert max(psi2.bond_dims()) <= 4
    with pytest.raises(ValueError):
        G.evolve_gate(psi, G.Gate("cx", (1, 3)))          # "Gate lanes must be contiguous" (Chain.jl:574)
