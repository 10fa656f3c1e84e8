"""GPU parity against the committed golden fixtures (tests/golden/golden_v1.json; generator and provenance in
tests/golden/make_golden.py): the CUDA path, called through the C-ABI on inputs regenerated from the seeds, must
reproduce the stored gauge-invariant numbers -- Schmidt values to 1e-12 sigma_1, scalars to 1e-10, kept counts,
contraction path and cut indices bit-exact -- and the analytic known answers stored with them."""
import json
import os

import numpy as np
import pytest

from oracle import chain as oc
from oracle import circuit as ocirc

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")))
SIG_TOL, OBS_TOL = 1e-12, 1e-10
Z = np.diag([1.0, -1.0]).astype(complex)


@pytest.fixture(scope="module")
def qb():
    import qrochet_b200 as q
    return q


@pytest.fixture(scope="module")
def ctx(qb):
    c = qb.Context(0)
    yield c
    c.close()


def c_(v):
    return complex(v[0], v[1])


def check_lams(got, want):
    assert len(got) == len(want)
    for b, (g, w) in enumerate(zip(got, want)):
        assert (g is None) == (w is None), b
        if w is not None:
            w = np.array(w)
            assert len(g) == len(w), b                                  # kept count bit-exact
            assert np.abs(g - w).max() <= SIG_TOL * w[0], b


def test_canonize_fixtures(qb, ctx):
    for c in GOLD["canonize"]:
        psi = qb.B200MPS(ctx, oc.rand_mps_arrays(np.random.default_rng(c["seed"]), c["n"], c["chi"])).canonize()
        check_lams(psi.lambdas(), c["lambdas"])
        assert abs(psi.expect([Z], [c["expect_Z_site"]])[0] - c_(c["expect_Z"])) <= OBS_TOL
        assert abs(psi.overlap(psi) - c_(c["overlap_self"])) <= OBS_TOL


def test_tebd_fixtures(qb, ctx):
    for c in GOLD["tebd"]:
        n, chi = c["n"], c["chi"]
        psi = qb.B200MPS(ctx, oc.rand_mps_arrays(np.random.default_rng(c["seed"]), n, chi)).canonize()
        rng = np.random.default_rng(c["seed"] + 1)
        gates = [np.reshape(oc.haar_unitary(rng), (2, 2, 2, 2), order="F") for _ in c["bonds"]]
        kept, _ = psi.evolve_circuit(gates, c["bonds"], maxdim=c["maxdim"], iscanonical=True, renormalize=True)
        assert kept == c["kept"]
        check_lams(psi.lambdas(), c["lambdas"])
        assert abs(psi.norm() - c["norm"]) <= OBS_TOL
        other = qb.B200MPS(ctx, oc.rand_mps_arrays(np.random.default_rng(c["seed"] + 2), n, chi))
        assert abs(psi.overlap(other) - c_(c["overlap_with_seed_plus_2"])) <= OBS_TOL


def test_mixed_canonize_and_mpo_fixtures(qb, ctx):
    for c in GOLD["mixed_canonize"]:
        psi = qb.B200MPS(ctx, oc.rand_mps_arrays(np.random.default_rng(c["seed"]), c["n"], c["chi"]))
        psi.mixed_canonize(c["center"])
        lam = psi.lambdas()[c["center"] - 2]
        w = np.array(c["lambda"])
        assert len(lam) == len(w) and np.abs(lam - w).max() <= SIG_TOL * w[0]
    for c in GOLD["mpo"]:
        psi = qb.B200MPS(ctx, oc.rand_mps_arrays(np.random.default_rng(c["seed"]), c["n"], c["chi"]))
        mpo = qb.heisenberg_mpo_arrays(c["n"])
        assert abs(psi.expect_mpo(mpo) - c_(c["expect_H"])) <= OBS_TOL * max(1.0, abs(c_(c["expect_H"])))
        psi.apply_mpo(mpo).compress(maxdim=c["maxdim"])
        check_lams(psi.lambdas(), c["lambdas"])
        assert abs(psi.norm() ** 2 - c["norm2_after"]) <= OBS_TOL * max(1.0, c["norm2_after"])


def test_circuit_fixtures(qb, ctx):
    for c in GOLD["circuit"]:
        n = c["n"]
        gates = qb.random_fsim_circuit(n, c["depth"])
        ket, bra = ocirc.random_product_state(n, 1), ocirc.random_product_state(n, 2)
        arrays, modes = qb.amplitude_network(n, gates, ket, bra)
        sc = qb.SlicedContraction(ctx, arrays, modes, c["max_elements"])
        assert sc.sliced_modes == c["sliced_modes"]                      # cut-index choice bit-exact
        assert [list(p) for p in sc.path] == c["path"]
        assert abs(sc.contract() - c_(c["amplitude"])) <= OBS_TOL * max(abs(c_(c["amplitude"])), 1e-3)


def test_analytic_known_answers(qb, ctx):
    a = GOLD["analytic"]
    H1 = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)
    cnot = np.zeros((4, 4), dtype=complex)        # control = first lane (fastest bit), target = second lane
    for q1 in (0, 1):
        for q2 in (0, 1):
            cnot[q1 + 2 * (q2 ^ q1), q1 + 2 * q2] = 1.0
    cnot = np.reshape(cnot, (2, 2, 2, 2), order="F")
    e0, e1 = np.array([1.0, 0.0]), np.array([0.0, 1.0])
    for n, key in ((2, "bell"), (5, "ghz5")):
        psi = qb.B200MPS.from_product(ctx, [e0] * n).canonize()
        psi.evolve(H1, [1])
        psi.evolve_circuit([cnot] * (n - 1), list(range(1, n)), iscanonical=True)
        for lam in psi.lambdas():
            assert len(lam) == 2 and np.abs(lam - np.array(a[key]["lambda"])).max() <= SIG_TOL
        zz = psi.expect([Z], [1])[0]
        assert abs(zz) <= OBS_TOL                                        # <Z_1> = 0 on Bell / GHZ
        assert abs(psi.norm() - 1.0) <= OBS_TOL
    s = 1 / np.sqrt(2)
    singlet = qb.B200MPS(ctx, [np.eye(2, dtype=complex), np.array([[0, -s], [s, 0]], dtype=complex)])
    assert abs(singlet.expect_mpo(qb.heisenberg_mpo_arrays(2)) - a["singlet_energy"]["value"]) <= OBS_TOL
    neel = qb.B200MPS.from_product(ctx, [e0 if k % 2 == 0 else e1 for k in range(8)])
    assert abs(neel.expect_mpo(qb.heisenberg_mpo_arrays(8)) - a["neel_energy_n8"]["value"]) <= OBS_TOL
