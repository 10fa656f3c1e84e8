"""CPU-side tests: the C-ABI library loads and exports every symbol the header declares, the host logic
(rand MPS generator, sweep order, circuit -> network builder) and the C++ planner against the oracle's
restatement (same path, same cut indices -- bit-exact)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import chain as oc
from oracle import circuit as ocirc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    import qrochet_b200 as qb
    header = open(os.path.join(ROOT, "include", "qrochet_b200.h")).read()
    declared = set(re.findall(r"\b(qb200_[a-z0-9_]+)\s*\(", header))
    lib = ctypes.CDLL(qb._capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(qb._capi.EXPORTS), declared ^ set(qb._capi.EXPORTS)


def test_no_cpu_fallback():
    import qrochet_b200 as qb
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(qb.QB200Error):
        qb.Context(0)


def test_rand_mps_generator_matches_reference_properties():
    import qrochet_b200 as qb
    n, chi = 8, 10
    arrays = qb.rand_mps_arrays(np.random.default_rng(0), n, chi)
    # Chain_test.jl:206-221: norm 1, all dims <= chi; Appendix B: right-canonical on sites 2..n
    q = oc.Chain(arrays)
    assert np.isclose(q.norm(), 1.0)
    assert max(max(a.shape) for a in arrays) <= chi
    for k in range(2, n + 1):
        assert q.isrightcanonical(oc.site(k))
    ref = oc.rand_mps_arrays(np.random.default_rng(0), n, chi)
    assert [a.shape for a in arrays] == [a.shape for a in ref]
    assert qb.bond_dims(64, 1024).count(1024) == 45 and qb.bond_dims(16, 32).count(32) == 7


def test_haar_gate_is_unitary_in_reference_layout():
    import qrochet_b200 as qb
    g = qb.haar_gate(np.random.default_rng(1))
    assert g.shape == (2, 2, 2, 2)
    u = np.reshape(g, (4, 4), order="F")
    assert np.allclose(u.conj().T @ u, np.eye(4))


def test_circuit_network_matches_oracle_builder():
    import qrochet_b200 as qb
    gates = qb.random_fsim_circuit(8, 3)
    ogates = ocirc.random_fsim_circuit(8, 3)
    assert [(a, b) for a, b, _ in gates] == [(a, b) for a, b, _ in ogates]
    arrays, modes = qb.amplitude_network(8, gates)
    oarrays, omodes = ocirc.amplitude_network(8, ogates)
    assert modes == omodes and all(np.array_equal(x, y) for x, y in zip(arrays, oarrays))
    assert len(arrays) == 12 + 16


@pytest.mark.parametrize("n,depth,maxel,optimizer", [(8, 3, 0, 1), (10, 4, 2 ** 6, 1), (12, 4, 2 ** 5, 1), (14, 6, 2 ** 7, 1),
                                                     (12, 4, 2 ** 5, 0), (16, 6, 2 ** 8, 0)])
def test_planner_matches_oracle_bit_exactly(n, depth, maxel, optimizer):
    """The C++ planner (csrc/tn_plan.cu: ContractSimplification + multi-start greedy + sub-tree reconfiguration +
    findslices) against the same rules restated in Python (oracle/circuit.py::plan): same contraction path step by
    step, same cut indices in the same order, same flop counts.  optimizer = 0 is the round-1 single greedy tree."""
    import qrochet_b200 as qb
    gates = qb.random_fsim_circuit(n, depth)
    ket, bra = ocirc.random_product_state(n, 5), ocirc.random_product_state(n, 6)
    arrays, modes = qb.amplitude_network(n, gates, ket, bra)
    plan = qb.SlicedContraction(None, arrays, modes, maxel, optimizer=optimizer)
    extents = {x: 2 for m in modes for x in m}
    want = ocirc.plan(modes, extents, maxel, optimizer=optimizer)
    assert plan.path == want["path"]
    assert plan.sliced_modes == want["sliced"]          # first-occurrence slice choice, bit-exact
    assert plan.nslices == want["nslices"] == 2 ** len(want["sliced"])
    assert plan.flops_per_slice == 8.0 * want["macs_per_slice"]
    assert plan.flops_invariant == 8.0 * want["macs_invariant"]
    if maxel:
        assert plan.max_intermediate <= maxel


def test_planner_local_search_beats_the_single_greedy_tree():
    """What the local search buys on a circuit of the benchmark's family (pure host work): far fewer flops, never more."""
    import qrochet_b200 as qb
    n, depth = 24, 6
    gates = qb.random_fsim_circuit(n, depth)
    arrays, modes = qb.amplitude_network(n, gates)
    old = qb.SlicedContraction(None, arrays, modes, 2 ** 14, optimizer=0)
    new = qb.SlicedContraction(None, arrays, modes, 2 ** 14, optimizer=1)
    total = lambda p: p.flops_per_slice * p.nslices + p.flops_invariant  # noqa: E731
    assert total(new) * 4 <= total(old)
    assert new.max_intermediate <= 2 ** 14 and old.max_intermediate <= 2 ** 14


def test_oracle_sliced_sum_equals_statevector():
    n, depth = 10, 4
    gates = ocirc.random_fsim_circuit(n, depth)
    assert abs(ocirc.statevector_amplitude(n, gates) - 1.0) < 1e-12   # FSim leaves |0..0> alone (the example's case)
    ket, bra = ocirc.random_product_state(n, 1), ocirc.random_product_state(n, 2)
    arrays, modes = ocirc.amplitude_network(n, gates, ket, bra)
    extents = {x: 2 for m in modes for x in m}
    exact = ocirc.statevector_amplitude(n, gates, ket, bra)
    full, nsl = ocirc.contract_sliced(arrays, modes, ocirc.plan(modes, extents, 0))
    assert nsl == 1 and abs(full - exact) < 1e-12
    pl = ocirc.plan(modes, extents, 2 ** 5)
    total, nsl = ocirc.contract_sliced(arrays, modes, pl)
    assert nsl > 1 and abs(total - exact) < 1e-12
    # every slice exactly once when dealt round-robin to 3 ranks
    parts = [ocirc.contract_sliced(arrays, modes, pl, r, 3)[0] for r in range(3)]
    assert abs(sum(parts) - exact) < 1e-12


def test_circuit_front_end_structure():
    """`Quantum(circuit)` structure as test/integration/Quac_test.jl:4-16 checks it for QFT(3): n inputs, n outputs,
    every open index is a site index; plus the SWAP-as-wire-relabel rule (QrochetYaoExt.jl:23-27)."""
    import qrochet_b200 as qb
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)

    def cphase(k):
        return np.diag([1, 1, 1, np.exp(2j * np.pi / 2 ** k)])
    # QFT(3): H0 CP(0,1) CP(0,2) H1 CP(1,2) H2 SWAP(0,2)
    gates = [((0,), h), ((0, 1), cphase(2)), ((0, 2), cphase(3)), ((1,), h), ((1, 2), cphase(2)), ((2,), h),
             ((0, 2), "swap")]
    arrays, modes, inputs, outputs = qb.circuit_network(3, gates)
    assert len(arrays) == 6 and len(inputs) == len(outputs) == 3
    count = {}
    for m in modes:
        for x in m:
            count[x] = count.get(x, 0) + 1
    opened = {x for x, c in count.items() if c == 1}
    assert opened == set(inputs) | set(outputs)
    # dense check: contracting the network gives the circuit unitary (with the reference's in/out labelling the
    # tensors are the transposed gates: symmetric here), SWAP applied as a relabelling
    letters = {x: i for i, x in enumerate(sorted(count))}
    args = []
    for a, m in zip(arrays, modes):
        args += [a, [letters[x] for x in m]]
    out = np.einsum(*args, [letters[x] for x in outputs] + [letters[x] for x in inputs])
    U = np.reshape(out, (8, 8), order="F")
    assert np.allclose(U.conj().T @ U, np.eye(8))
    # known answer: state-vector application of the gate list.  Two quirks of the reference are replicated knowingly
    # (SURVEY.md Appendix A): array dims 1..k are labelled as the incoming wires (every tensor is the transposed gate:
    # all gates here are symmetric), and a SWAP exchanges the two WHOLE wires (QrochetYaoExt.jl:23-27,
    # QrochetQuacExt.jl:26-30) -- inputs included -- so a trailing SWAP yields P U P, not P U.
    from oracle import statevector as sv

    def swap02(psi):
        return np.reshape(np.swapaxes(np.reshape(psi, (2, 2, 2), order="F"), 0, 2), -1, order="F")
    for col in range(8):
        psi = np.zeros(8, complex)
        psi[col] = 1.0
        psi = swap02(psi)
        for qubits, mat in gates[:-1]:
            psi = sv.apply_gate(psi, np.asarray(mat).T, [q + 1 for q in qubits], 3)
        assert np.allclose(U[:, col], swap02(psi))


# ---- the device Chain's bookkeeping (site map, bonds, boundary, socket) is host logic: run the reference's
# ---- constructor tests (Chain_test.jl:2-188) against it with a stand-in for the device array type
class _HostArray:
    """Carries only what the bookkeeping reads from a DeviceArray: shape / ndim / dtype."""

    def __init__(self, a):
        self.shape, self.ndim, self.dtype = tuple(a.shape), a.ndim, 1 if a.dtype == np.complex64 else 0


class _HostCtx:
    def array(self, a):
        return _HostArray(np.asarray(a))


import _chain_ctor_cases as ctor  # noqa: E402


@pytest.mark.parametrize("case", ctor.ALL, ids=lambda f: f.__name__)
def test_device_chain_constructors_host_logic(case):
    import qrochet_b200 as qb
    case(lambda arrays, **kw: qb.chain.Chain(_HostCtx(), arrays, **kw), qb.chain.site)


def test_rand_mpo_generator_matches_reference_properties():
    """`rand(Chain, Open, Operator)` (Chain.jl:260-297, Chain_test.jl:223-235): Frobenius norm 1, bonds <= chi,
    shapes identical to the oracle's restatement (which orthonormalises with Gram-Schmidt like the reference)."""
    import qrochet_b200 as qb
    for n, chi in [(8, 10), (5, 3), (7, 100)]:
        arrays = qb.rand_mpo_arrays(np.random.default_rng(0), n, chi)
        ref = oc.rand_mpo_arrays(np.random.default_rng(0), n, chi)
        assert [a.shape for a in arrays] == [a.shape for a in ref]
        q = oc.Chain(arrays, socket="operator")
        assert np.isclose(q.norm(), 1.0)
        assert max(d for a in arrays for d in a.shape[2:]) <= chi


@pytest.mark.parametrize("steps,warmup", [(1, 0), (3, 1)])
def test_bench_reference_arm_line_has_the_contract_keys(steps, warmup):
    """`bench.py --impl reference` (the CPU arm the driver runs beside the B200 arm) prints ONE JSON line with the
    contract's keys -- as one real full sweep (--steps 1 --warmup 0) and as the bounded sample (one bulk-bond evolve!
    per step) -- never maps the product library, and reports a step time that is the step's own wall time.  Run here
    on a tiny chain so the whole CPU suite stays fast."""
    import json
    import subprocess
    import sys
    code = ("import sys, runpy; sys.argv = ['bench.py'] + %r; runpy.run_path(%r, run_name='__main__'); "
            "print('MAPPED', any('qrochet_b200' in l for l in open('/proc/self/maps')))")
    argv = ["--impl", "reference", "--steps", str(steps), "--warmup", str(warmup), "--sites", "8", "--bond-dim", "8"]
    t0 = __import__("time").perf_counter()
    out = subprocess.run([sys.executable, "-c", code % (argv, os.path.join(ROOT, "bench.py"))], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    wall = __import__("time").perf_counter() - t0
    assert out.returncode == 0, out.stderr[-2000:]
    assert "MAPPED False" in out.stdout                     # the reference arm must not load libqrochet_b200.so
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "sweeps/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["ms_per_step"] * d["steps"] * 1e-3 <= wall      # the timed region fits in the run (no extrapolated time)
    import bench
    assert d["config"] == json.loads(json.dumps(bench.config_dict(8, 8)))   # same `config` object as the B200 arm


def test_product_ansatz_host_mirror_matches_oracle():
    """qrochet_b200.Product (Product.jl restated on the host) against the oracle's restatement and the reference's
    own test (test/Ansatz/Product_test.jl)."""
    import qrochet_b200 as qb
    from oracle.chain import Product as OProduct
    r = np.random.default_rng(31)
    vecs = [r.random(2) + 1j * r.random(2) for _ in range(3)]
    mats = [r.random((2, 2)) for _ in range(3)]
    p, o = qb.Product(vecs), OProduct(vecs)
    assert p.socket == "state" and np.isclose(p.norm(), o.norm())
    assert np.isclose(p.copy().normalize_().norm(), 1.0)
    pm, om = qb.Product(mats), OProduct(mats)
    assert pm.socket == "operator" and np.isclose(pm.norm(), om.norm()) and np.isclose(pm.opnorm(), om.opnorm())
    assert np.isclose(pm.copy().normalize_().norm(), 1.0)
    q = qb.Product([r.random(2) + 1j * r.random(2) for _ in range(3)])
    assert np.isclose(p.overlap(q), o.overlap(OProduct(q.arrays)))
    z, one = qb.Product.zeros(4), qb.Product.ones(4)
    assert z.overlap(z) == 1 and z.overlap(one) == 0 and [a.tolist() for a in one.arrays] == [[False, True]] * 4
    with pytest.raises(TypeError):
        qb.Product([np.zeros((2, 2, 2))])
