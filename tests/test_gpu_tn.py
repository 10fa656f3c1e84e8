"""GPU parity of the sliced circuit-amplitude contraction (examples/distributed.jl pattern) through the C-ABI:
sum over all slices == unsliced contraction == dense state-vector amplitude; same path and cut indices as the
oracle's planner; round-robin slice -> rank dealing covers every slice exactly once."""
import numpy as np
import pytest

from oracle import circuit as ocirc

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


@pytest.mark.parametrize("n,depth,maxel", [(6, 2, 0), (10, 4, 0), (10, 4, 2 ** 5), (14, 5, 2 ** 7), (18, 6, 2 ** 10)])
def test_sliced_amplitude_matches_statevector_and_oracle(qb, ctx, n, depth, maxel):
    gates = qb.random_fsim_circuit(n, depth)
    ket, bra = ocirc.random_product_state(n, 1), ocirc.random_product_state(n, 2)
    arrays, modes = qb.amplitude_network(n, gates, ket, bra)
    sc = qb.SlicedContraction(ctx, arrays, modes, maxel)
    exact = ocirc.statevector_amplitude(n, ocirc.random_fsim_circuit(n, depth), ket, bra)
    assert 1e-6 < abs(exact) < 0.99          # a non-trivial amplitude (<0|U|0> of an FSim circuit is exactly 1)
    extents = {x: 2 for m in modes for x in m}
    want_plan = ocirc.plan(modes, extents, maxel)
    assert sc.path == want_plan["path"] and sc.sliced_modes == want_plan["sliced"]
    got = sc.contract()
    assert abs(got - exact) <= 1e-10 * max(abs(exact), 1e-3)
    if maxel:
        assert sc.nslices > 1 and sc.max_intermediate <= maxel
        # slice s on rank s mod W: the partial sums of 3 "ranks" add up to the same amplitude
        parts = [sc.contract(first_slice=r, stride=3) for r in range(3)]
        assert abs(sum(parts) - exact) <= 1e-10 * max(abs(exact), 1e-3)
        opart, _ = ocirc.contract_sliced(arrays, modes, want_plan, 1, 3)
        assert abs(parts[1] - opart) <= 1e-10 * max(abs(opart), 1e-3)


def test_hyperindex_network(qb, ctx):
    """A network with an index shared by three tensors (Tenet hyper-index, as Λ on a bond)."""
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 4)) + 1j * rng.standard_normal((3, 4))
    lam = rng.random(4) + 0j
    b = rng.standard_normal((4, 3)) + 1j * rng.standard_normal((4, 3))
    sc = qb.SlicedContraction(ctx, [a, lam, b], [(0, 1), (1,), (1, 0)], 0)
    want = np.einsum("ij,j,ji->", a, lam, b)
    assert abs(sc.contract() - want) < 1e-12 * abs(want)
