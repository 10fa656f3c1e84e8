"""Generates tests/golden/golden_v1.json: seeded inputs -> gauge-invariant outputs of the hot path.

The reference ships NO golden vectors and cannot run here (no Julia toolchain), so these fixtures are
produced by the CPU oracle (`oracle/`, the restatement of Chain.jl / examples/distributed.jl) and cross-checked inside
this script against dense state-vector arithmetic before they are written.  They (a) freeze the oracle against
regressions and (b) are what the `-m gpu` golden test compares the CUDA path with.  Inputs are regenerated from the
seeds by the tests (default_rng is stable across NumPy versions by policy).

    python tests/golden/make_golden.py        # rewrites golden_v1.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import chain as oc  # noqa: E402
from oracle import circuit as ocirc  # noqa: E402
from oracle.chain import site  # noqa: E402

Z = np.diag([1.0, -1.0]).astype(complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)


def cplx(z):
    return [float(np.real(z)), float(np.imag(z))]


def lam_list(lams):
    return [None if x is None else [float(v) for v in x] for x in lams]


def case_canonize(seed, n, chi):
    arrays = oc.rand_mps_arrays(np.random.default_rng(seed), n, chi)
    psi = oc.Chain(arrays)
    dense = psi.to_dense()
    psi.canonize()
    assert np.allclose(psi.to_dense(), dense, atol=1e-12)            # canonize! only changes the gauge
    s = n // 2
    ez = psi.expect([oc.Dense(Z, [site(s), site(s, True)])])
    want = np.vdot(dense, np.kron(np.eye(2 ** (n - s)), np.kron(Z, np.eye(2 ** (s - 1)))) @ dense)
    assert abs(ez - want) < 1e-12
    return {"seed": seed, "n": n, "chi": chi, "lambdas": lam_list(psi.lambdas()), "expect_Z_site": s,
            "expect_Z": cplx(ez), "overlap_self": cplx(psi.overlap(psi))}


def case_tebd(seed, n, chi, maxdim, bonds):
    arrays = oc.rand_mps_arrays(np.random.default_rng(seed), n, chi)
    psi = oc.Chain(arrays)
    psi.canonize()
    rng = np.random.default_rng(seed + 1)
    kept, dws = [], []
    for b in bonds:
        U = oc.haar_unitary(rng)
        before = psi.lambdas()[b - 1]
        th_dim = None
        psi.evolve(oc.gate(U, [b, b + 1]), iscanonical=True, maxdim=maxdim, renormalize=True)
        kept.append(len(psi.lambdas()[b - 1]))
    other = oc.Chain(oc.rand_mps_arrays(np.random.default_rng(seed + 2), n, chi))
    return {"seed": seed, "n": n, "chi": chi, "maxdim": maxdim, "bonds": bonds, "kept": kept,
            "lambdas": lam_list(psi.lambdas()), "norm": float(psi.norm()),
            "overlap_with_seed_plus_2": cplx(psi.overlap(other))}


def case_mixed(seed, n, chi, center):
    psi = oc.Chain(oc.rand_mps_arrays(np.random.default_rng(seed), n, chi))
    psi.mixed_canonize(site(center))
    lam = psi.lambda_between(site(center - 1), site(center)).data
    return {"seed": seed, "n": n, "chi": chi, "center": center, "lambda": [float(v) for v in lam]}


def case_mpo(seed, n, chi, maxdim):
    arrays = oc.rand_mps_arrays(np.random.default_rng(seed), n, chi)
    mpo = oc.heisenberg_mpo_arrays(n)
    psi = oc.Chain(arrays)
    e = oc.expect_mpo(psi, mpo)
    dense = psi.to_dense()
    H = oc.mpo_to_dense(mpo)
    assert abs(e - np.vdot(dense, H @ dense)) < 1e-11
    phi = oc.compress(oc.Chain(oc.apply_mpo_arrays(arrays, mpo)), maxdim=maxdim)
    return {"seed": seed, "n": n, "chi": chi, "maxdim": maxdim, "expect_H": cplx(e), "lambdas": lam_list(phi.lambdas()),
            "norm2_after": float(phi.norm() ** 2)}


def case_circuit(n, depth, maxel):
    gates = ocirc.random_fsim_circuit(n, depth)
    ket, bra = ocirc.random_product_state(n, 1), ocirc.random_product_state(n, 2)
    arrays, modes = ocirc.amplitude_network(n, gates, ket, bra)
    extents = {x: 2 for m in modes for x in m}
    pl = ocirc.plan(modes, extents, maxel)
    amp, _ = ocirc.contract_sliced(arrays, modes, pl)
    exact = ocirc.statevector_amplitude(n, gates, ket, bra)
    assert abs(amp - exact) < 1e-12
    return {"n": n, "depth": depth, "max_elements": maxel, "sliced_modes": [int(x) for x in pl["sliced"]],
            "path": [[int(a), int(b)] for a, b in pl["path"]], "amplitude": cplx(amp)}


def analytic():
    """Known answers that need no oracle at all."""
    s = 1 / np.sqrt(2)
    return {
        "bell": {"doc": "CNOT (H x 1)|00>: Schmidt values (1/sqrt2, 1/sqrt2), <Z1> = 0, <Z1 Z2> = 1", "lambda": [s, s]},
        "ghz5": {"doc": "5-qubit GHZ by a CNOT ladder: every bond has Schmidt values (1/sqrt2, 1/sqrt2)", "lambda": [s, s]},
        "singlet_energy": {"doc": "Heisenberg H = S1.S2 on the singlet: -3/4", "value": -0.75},
        "neel_energy_n8": {"doc": "<Neel|H|Neel> for the open n=8 Heisenberg chain: -(n-1)/4", "value": -7 / 4},
    }


if __name__ == "__main__":
    out = {
        "doc": "gauge-invariant outputs of the oracle on seeded inputs; tolerances: sigma 1e-12 sigma_1, scalars 1e-10",
        "canonize": [case_canonize(101, 8, 8), case_canonize(102, 10, 12)],
        "tebd": [case_tebd(201, 8, 8, 6, [1, 3, 5, 7, 2, 4, 6, 4, 4, 1]), case_tebd(204, 10, 16, 16, [5, 4, 6, 3, 7, 5, 9, 1])],
        "mixed_canonize": [case_mixed(301, 9, 8, 5), case_mixed(302, 8, 6, 2)],
        "mpo": [case_mpo(401, 8, 6, 10)],
        "circuit": [case_circuit(10, 4, 2 ** 5), case_circuit(12, 5, 2 ** 4), case_circuit(14, 6, 2 ** 6)],
        "analytic": analytic(),
    }
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")
