"""GPU parity AT THE BENCHMARKED SIZES (BASELINE.json configs 2, 3, 4): the device path and the CPU
oracle run the same workload from the same state and every gauge-invariant quantity is compared -- all kept counts
bit-exact, every Schmidt value to 1e-12 sigma_1, overlaps / <H> / norms to 1e-10 (north star).

The oracle is started from the device's canonized state (`oracle.chain.chain_from_vidal`): `canonize!` itself is
compared with the oracle at n = 16 / chi = 32 (test_gpu_mps.py) and at chi = 1024 against its defining properties
(test_gpu_fullsize.py); re-running it on the CPU at n = 64, chi = 1024 would add minutes of LAPACK time and no
information.  CPU cost of this file on the GPU box's host cores: about 4 minutes, almost all of it zgesdd."""
import time

import numpy as np
import pytest

from oracle import chain as oc
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


def sweep_bonds(n):
    return list(range(1, n, 2)) + list(range(2, n, 2))     # odd bonds, then even bonds (bench.py)


def haar(layer, bond):
    return oc.haar_unitary(np.random.default_rng(2000 + layer * 64 + bond))   # the bench's gate stream


def to_device(qb, ctx, o: oc.Chain, n):
    """Upload the oracle's Vidal chain (site tensors in whatever index order the oracle left them)."""
    sites = []
    for k in range(1, n + 1):
        t = o.tensor_at(site(k))
        order = ([o.leftindex(site(k))] if k > 1 else []) + [o.sites[site(k)]] + ([o.rightindex(site(k))] if k < n else [])
        a = t.permute(order).data
        if k == 1:
            a = a[None, ...]
        if k == n:
            a = a[..., None]
        sites.append(np.asfortranarray(a))
    return qb.B200MPS.from_sites(ctx, sites, o.lambdas(), form=1)


def compare_states(qb, ctx, g, o, n):
    """Schmidt values bond by bond, then the state itself: |<oracle|device>| = |oracle| |device| (gauge invariant)."""
    gl, ol = g.lambdas(), o.lambdas()
    for b in range(n - 1):
        assert len(gl[b]) == len(ol[b]), f"bond {b + 1}: kept {len(gl[b])} vs oracle {len(ol[b])}"
        assert np.abs(gl[b] - ol[b]).max() <= SIG_TOL * ol[b][0], f"bond {b + 1}"
    og = to_device(qb, ctx, o, n)
    ng, no, ov = g.norm(), og.norm(), g.overlap(og)
    assert abs(ng - no) <= OBS_TOL * no
    assert abs(abs(ov) - ng * no) <= OBS_TOL * ng * no
    assert abs(ov.imag) <= OBS_TOL * ng * no and ov.real > 0       # same gauge-fixed sign conventions are NOT assumed
    return ng


def run_tebd(qb, ctx, n, chi, layers):
    arrays = qb.rand_mps_arrays(np.random.default_rng(1000 + 4), n, chi)
    g = qb.B200MPS(ctx, arrays).canonize()
    del arrays
    o = oc.chain_from_vidal([g.site(s) for s in range(n)], g.lambdas())
    order = sweep_bonds(n)
    kept_g, kept_o = [], []
    t_gpu = t_cpu = 0.0
    for layer in range(layers):
        mats = [haar(layer, b) for b in order]
        t0 = time.perf_counter()
        kept, dw = g.evolve_circuit([np.reshape(u, (2, 2, 2, 2), order="F") for u in mats], order, maxdim=chi,
                                    iscanonical=True, renormalize=True)
        t_gpu += time.perf_counter() - t0
        kept_g += kept
        t0 = time.perf_counter()
        for u, b in zip(mats, order):
            o.evolve(oc.gate(u, [b, b + 1]), iscanonical=True, maxdim=chi, renormalize=True)
            kept_o.append(len(o.lambdas()[b - 1]))
        t_cpu += time.perf_counter() - t0
    assert kept_g == kept_o                                   # every truncation decision, bit-exact
    norm = compare_states(qb, ctx, g, o, n)
    print(f"\n[n={n} chi={chi} x{layers}] device {t_gpu:.1f} s, oracle {t_cpu:.1f} s, |psi| after = {norm:.6f}")
    return g, o


def test_config4_full_tebd_sweep_matches_oracle(qb, ctx):
    """C4, the headline: ONE full sweep (63 evolve! calls, odd then even bonds) at n = 64, chi = 1024 on the device and on
    the oracle from the same Vidal state and the same Haar gates (Chain.jl:606-722)."""
    n, chi = 64, 1024
    g, o = run_tebd(qb, ctx, n, chi, layers=1)
    assert g.bond_dims() == qb.bond_dims(n, chi)
    # renormalize=True normalises every touched Schmidt vector (Chain.jl:653-654)
    for lam in g.lambdas():
        assert abs(np.sum(lam ** 2) - 1.0) <= OBS_TOL


def test_config2_brickwork_layers_match_oracle(qb, ctx):
    """C2: n = 64, chi = 256, 4 brickwork layers of random two-site gates truncated to chi_max = 256."""
    run_tebd(qb, ctx, 64, 256, layers=4)


def test_config3_mixed_canonize_and_mpo_expectation_match_oracle(qb, ctx):
    """C3 (first half): n = 64, chi = 512, mixed_canonize!(psi, Site(32)) (Chain.jl:509-524) and <psi|H|psi> with the
    Heisenberg MPO (D = 5)."""
    n, chi = 64, 512
    arrays = qb.rand_mps_arrays(np.random.default_rng(1000 + 3), n, chi)
    g = qb.B200MPS(ctx, arrays)
    o = oc.Chain(arrays)
    g.mixed_canonize(32)
    o.mixed_canonize(site(32))
    lam_g, lam_o = g.lambdas()[30], o.lambdas()[30]            # the centre SVD leaves Λ on bond (31, 32)
    assert len(lam_g) == len(lam_o) == chi
    assert np.abs(lam_g - lam_o).max() <= SIG_TOL * lam_o[0]
    assert all(l is None for k, l in enumerate(g.lambdas()) if k != 30)
    mpo = qb.heisenberg_mpo_arrays(n)
    want = oc.expect_mpo(o, mpo)
    got = g.expect_mpo(mpo)
    assert abs(got - want) <= OBS_TOL * abs(want)
    assert abs(g.norm() - o.norm()) <= OBS_TOL


def test_config3_mpo_application_with_truncation_matches_oracle(qb, ctx):
    """C3 (second half): H|psi> and its compression back to chi = 512 (bonds chi*D = 2560 -> 512).  n = 24 keeps the
    full chi*D = 2560 bulk shape on 7 bonds while the oracle's QR / SVD sweeps stay under a minute."""
    n, chi = 24, 512
    arrays = qb.rand_mps_arrays(np.random.default_rng(1000 + 3), n, chi)
    mpo = qb.heisenberg_mpo_arrays(n)
    g = qb.B200MPS(ctx, arrays).apply_mpo(mpo)
    assert max(g.bond_dims()) == chi * 5
    g.compress(maxdim=chi)
    o = oc.compress(oc.Chain(oc.apply_mpo_arrays(arrays, mpo)), maxdim=chi)
    assert g.bond_dims() == [len(l) for l in o.lambdas()]
    assert g.bond_dims().count(chi) == 7
    compare_states(qb, ctx, g, o, n)
    # <psi| (H psi)_compressed> against the oracle's
    psi = qb.B200MPS(ctx, arrays)
    want = o.overlap(oc.Chain(arrays))
    got = g.overlap(psi)
    assert abs(got - want) <= OBS_TOL * abs(want)
