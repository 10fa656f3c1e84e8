"""Host-side circuit front-end (ext/QrochetQuacExt.jl restated: `Dense(::Gate)` :7-13, `Quantum(::Circuit)` :17-48) and
the gate-list text format: gate-table identities, the reference's own Quac test (test/integration/Quac_test.jl:4-16:
QFT(3) has n inputs, n outputs, every open index is a site index) and a file round trip."""
import numpy as np
import pytest

import qrochet_b200 as qb
from qrochet_b200 import gates as G
from oracle import statevector as sv


def test_gate_table_is_unitary_and_consistent():
    rng = np.random.default_rng(0)
    for name, (nl, npar, _) in G.GATES.items():
        g = G.Gate(name, tuple(range(1, nl + 1)), tuple(rng.random(npar)))
        m = g.matrix()
        assert m.shape == (2 ** nl, 2 ** nl)
        assert np.allclose(m.conj().T @ m, np.eye(2 ** nl), atol=1e-14), name
        arr, sites = G.dense(g)
        assert arr.shape == (2,) * (2 * nl)
        assert sites == [(l, False) for l in g.lanes] + [(l, True) for l in g.lanes]      # QrochetQuacExt.jl:11
    H, X, Z = (G.Gate(n, (1,)).matrix() for n in "hxz")
    assert np.allclose(H @ Z @ H, X)
    assert np.allclose(G.Gate("s", (1,)).matrix() @ G.Gate("sd", (1,)).matrix(), np.eye(2))
    assert np.allclose(G.Gate("t", (1,)).matrix() @ G.Gate("t", (1,)).matrix(), G.Gate("s", (1,)).matrix())
    th = 0.37
    assert np.allclose(G.Gate("rz", (1,), (th,)).matrix(), np.diag([np.exp(-0.5j * th), np.exp(0.5j * th)]))
    # control = first lane = fastest bit: |10> (lane 1 set) -> |11>
    cx = G.Gate("cx", (1, 2)).matrix()
    e = np.zeros(4)
    e[1] = 1.0
    assert np.allclose(cx @ e, np.eye(4)[3])
    assert np.allclose(G.Gate("crz", (1, 2), (th,)).matrix()[1::2, 1::2], G.Gate("rz", (1,), (th,)).matrix())
    assert np.allclose(G.Gate("fsim", (1, 2), (0.3, 0.7)).matrix(), qb.fsim(0.3, 0.7))
    with pytest.raises(ValueError):
        G.Gate("cx", (1, 1))
    with pytest.raises(ValueError):
        G.Gate("rz", (1,))
    with pytest.raises(ValueError):
        G.Gate("nope", (1,))


def test_gate_list_text_format_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    u = np.linalg.qr(rng.standard_normal((4, 4)) + 1j * rng.standard_normal((4, 4)))[0]
    gates = [G.Gate("h", (1,)), G.Gate("cx", (1, 2)), G.Gate("rz", (3,), (0.25,)), G.Gate("fsim", (2, 3), (0.3, 0.7)),
             G.Gate("swap", (1, 3)), G.Gate("u", (2, 1), (), u)]
    path = tmp_path / "circuit.gates"
    G.dump(path, 3, gates)
    n, back = G.load(path)
    assert n == 3 and len(back) == len(gates)
    for a, b in zip(gates, back):
        assert a.name == b.name and a.lanes == b.lanes and np.array_equal(a.matrix(), b.matrix())   # repr() round-trips
    text = "qubits 2\n# Bell pair\nh 1\ncx 1 2   # entangle\n"
    n, bell = G.loads(text)
    psi = sv.zero_state(2)
    for g in bell:
        psi = sv.apply_gate(psi, g.matrix(), list(g.lanes), 2)
    assert np.allclose(psi, np.array([1, 0, 0, 1]) / np.sqrt(2))
    for bad in ("h 1\n", "qubits 2\ncx 1\n", "qubits 2\nh 3\n", "qubits 2\nfoo 1\n", "qubits 1\nu 1 1 1.0\n"):
        with pytest.raises(ValueError):
            G.loads(bad)


def test_quantum_of_a_circuit_is_the_qft():
    """test/integration/Quac_test.jl:4-16 on QFT(3), plus the dense known answer: the network contracts to the DFT
    matrix (with the reference's SWAP-as-wire-relabel rule the bit reversal acts on both ends)."""
    n = 3
    circ = G.qft(n)
    arrays, modes, inputs, outputs = G.circuit_to_network(n, circ)
    assert len(inputs) == len(outputs) == n
    count = {}
    for m in modes:
        for x in m:
            count[x] = count.get(x, 0) + 1
    assert {x for x, c in count.items() if c == 1} == set(inputs) | set(outputs)     # all open indices are sites
    # state-vector known answer: QFT|x> = DFT column (all gates symmetric, so the reference's transposed labelling is
    # invisible); SWAPs applied as gates here
    N = 2 ** n
    dft = np.array([[np.exp(2j * np.pi * a * b / N) for b in range(N)] for a in range(N)]) / np.sqrt(N)
    rev = [int(format(x, f"0{n}b")[::-1], 2) for x in range(N)]
    for col in range(N):
        psi = np.zeros(N, complex)
        psi[col] = 1.0
        for g in circ:
            psi = sv.apply_gate(psi, g.matrix(), list(g.lanes), n)
        # lane 1 is the fastest bit: our integer labels are bit-reversed with respect to the textbook's
        assert np.allclose(psi[rev], dft[:, rev[col]], atol=1e-12) or np.allclose(psi, dft[:, col], atol=1e-12)
