"""Circuit front-end of the reference (`ext/QrochetQuacExt.jl`): `Dense(::Gate)` (:7-13), `evolve!(::Ansatz, ::Gate)`
(:15) and `Quantum(::Circuit)` (:17-48; the network builder itself is `tn.circuit_network`), plus a plain-text gate-list
format so that circuits can be stored next to the benchmark inputs.

Quac itself is an un-vendored dependency of the reference (`Project.toml` weakdeps): the gate table below restates its
standard gate set [ext, from memory of Quac.jl's `Gate` types] with the usual matrices.  Conventions: lanes are 1-based;
the first lane of a gate is the FASTEST bit of its matrix, so the gate array of `Dense` is
`reshape(matrix, (2,)*2k)` column-major with dims (o_1..o_k, i_1..i_k) for sites [lanes..., lanes'...]
(`Dense.jl:21-34`)."""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass, field

import numpy as np

_S2 = 1.0 / math.sqrt(2.0)
_I2 = np.eye(2, dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.diag([1.0, -1.0]).astype(np.complex128)


def _rot(axis, theta):
    return math.cos(theta / 2) * _I2 - 1j * math.sin(theta / 2) * axis


def _kron_first_fastest(a, b):
    """Two-lane operator a (x) b with lane 1 (a) the fastest bit."""
    return np.kron(b, a)


def _controlled(u):
    """Control on the first lane (fastest bit), target on the second."""
    m = np.eye(4, dtype=np.complex128)
    for t_out in range(2):
        for t_in in range(2):
            m[1 + 2 * t_out, 1 + 2 * t_in] = u[t_out, t_in]
    return m


def _fsim(theta, phi):
    c, s = math.cos(theta), math.sin(theta)
    return np.array([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [0, 0, 0, cmath.exp(-1j * phi)]],
                    dtype=np.complex128)


def _u3(theta, phi, lam):
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array([[c, -cmath.exp(1j * lam) * s], [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]],
                    dtype=np.complex128)


# name -> (number of lanes, number of real parameters, matrix builder)
GATES = {
    "i": (1, 0, lambda: _I2.copy()),
    "x": (1, 0, lambda: _X.copy()),
    "y": (1, 0, lambda: _Y.copy()),
    "z": (1, 0, lambda: _Z.copy()),
    "h": (1, 0, lambda: _S2 * np.array([[1, 1], [1, -1]], dtype=np.complex128)),
    "s": (1, 0, lambda: np.diag([1, 1j]).astype(np.complex128)),
    "sd": (1, 0, lambda: np.diag([1, -1j]).astype(np.complex128)),
    "t": (1, 0, lambda: np.diag([1, cmath.exp(0.25j * math.pi)]).astype(np.complex128)),
    "td": (1, 0, lambda: np.diag([1, cmath.exp(-0.25j * math.pi)]).astype(np.complex128)),
    "rx": (1, 1, lambda th: _rot(_X, th)),
    "ry": (1, 1, lambda th: _rot(_Y, th)),
    "rz": (1, 1, lambda th: _rot(_Z, th)),
    "u2": (1, 2, lambda phi, lam: _u3(math.pi / 2, phi, lam)),
    "u3": (1, 3, _u3),
    "cx": (2, 0, lambda: _controlled(_X)),
    "cy": (2, 0, lambda: _controlled(_Y)),
    "cz": (2, 0, lambda: _controlled(_Z)),
    "crx": (2, 1, lambda th: _controlled(_rot(_X, th))),
    "cry": (2, 1, lambda th: _controlled(_rot(_Y, th))),
    "crz": (2, 1, lambda th: _controlled(_rot(_Z, th))),
    "cphase": (2, 1, lambda th: np.diag([1, 1, 1, cmath.exp(1j * th)]).astype(np.complex128)),
    "rxx": (2, 1, lambda th: math.cos(th / 2) * np.eye(4) - 1j * math.sin(th / 2) * _kron_first_fastest(_X, _X)),
    "ryy": (2, 1, lambda th: math.cos(th / 2) * np.eye(4) - 1j * math.sin(th / 2) * _kron_first_fastest(_Y, _Y)),
    "rzz": (2, 1, lambda th: math.cos(th / 2) * np.eye(4) - 1j * math.sin(th / 2) * _kron_first_fastest(_Z, _Z)),
    "fsim": (2, 2, _fsim),
    "swap": (2, 0, lambda: np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)),
    # "u": an explicit 2^k x 2^k matrix on k lanes (Quac's SU{N}); its entries are the parameters
}


@dataclass
class Gate:
    """`Quac.Gate`: a named operation on 1-based lanes with real parameters (or an explicit matrix for name "u")."""
    name: str
    lanes: tuple
    params: tuple = ()
    matrix_: np.ndarray | None = field(default=None, repr=False)

    def __post_init__(self):
        self.name = self.name.lower()
        self.lanes = tuple(int(l) for l in self.lanes)
        self.params = tuple(float(p) for p in self.params)
        if len(set(self.lanes)) != len(self.lanes) or any(l < 1 for l in self.lanes):
            raise ValueError(f"invalid lanes {self.lanes}")
        if self.name == "u":
            k = len(self.lanes)
            m = np.asarray(self.matrix_, dtype=np.complex128)
            if m.shape != (2 ** k, 2 ** k):
                raise ValueError(f"gate 'u' on {k} lanes needs a {2 ** k} x {2 ** k} matrix")
            self.matrix_ = m
            return
        if self.name not in GATES:
            raise ValueError(f"unknown gate '{self.name}'")
        nl, npar, _ = GATES[self.name]
        if len(self.lanes) != nl or len(self.params) != npar:
            raise ValueError(f"gate '{self.name}' takes {nl} lane(s) and {npar} parameter(s)")

    def matrix(self) -> np.ndarray:
        """2^k x 2^k matrix, first lane = fastest bit (`Matrix(gate)`)."""
        if self.name == "u":
            return self.matrix_
        return np.asarray(GATES[self.name][2](*self.params), dtype=np.complex128)


def dense(gate: Gate):
    """`Qrochet.Dense(gate::Gate)` (QrochetQuacExt.jl:7-13): (array, sites) with array dims (o_1..o_k, i_1..i_k) for
    sites [Site.(lanes)..., Site.(lanes; dual = true)...]; a site is (lane, dual)."""
    k = len(gate.lanes)
    arr = np.reshape(gate.matrix(), (2,) * (2 * k), order="F")
    sites = [(l, False) for l in gate.lanes] + [(l, True) for l in gate.lanes]
    return arr, sites


def evolve_gate(psi, gate: Gate, **kwargs):
    """`evolve!(qtn::Ansatz, gate::Gate; kwargs...) = evolve!(qtn, Dense(gate); kwargs...)` (QrochetQuacExt.jl:15) for
    any host mirror with an `evolve(array, lanes, **kwargs)` method (B200MPS, chain.Chain)."""
    arr, _ = dense(gate)
    return psi.evolve(arr, list(gate.lanes), **kwargs)


def evolve_gates(psi, gates, **kwargs):
    """A circuit on a chain: the loop `for g in circuit evolve!(psi, g; kwargs...)`.  Maximal runs of two-lane gates go
    to the device as ONE gate list (`evolve_circuit`: dependency-scheduled, identical results to the loop); SWAPs are
    applied as gates (a chain has no wires to relabel).  Lanes must be contiguous, as in the reference (Chain.jl:574)."""
    run_g, run_b, out = [], [], []

    def flush():
        if run_g:
            out.append(psi.evolve_circuit(run_g, run_b, **kwargs))
            run_g.clear()
            run_b.clear()

    for g in gates:
        arr, _ = dense(g)
        if len(g.lanes) == 2 and hasattr(psi, "evolve_circuit"):
            a, b = g.lanes
            if abs(a - b) != 1:
                raise ValueError("Gate lanes must be contiguous")
            if a > b:
                arr = np.transpose(arr, (1, 0, 3, 2))
                a = b
            run_g.append(np.asfortranarray(arr))
            run_b.append(a)
        else:
            flush()
            out.append(psi.evolve(arr, list(g.lanes), **kwargs))
    flush()
    return out


def circuit_to_network(n, gates):
    """`Quantum(circuit)` (QrochetQuacExt.jl:17-48): (arrays, modes, inputs, outputs) through `tn.circuit_network`; SWAP
    exchanges the wires (:26-30)."""
    from .tn import circuit_network

    return circuit_network(n, [(tuple(l - 1 for l in g.lanes), None if g.name == "swap" else g.matrix()) for g in gates])


# ---- gate-list text format -------------------------------------------------------------------------------------------
# One gate per line: `name lane [lane ...] [param ...]`, lanes 1-based integers, parameters floats (radians), `#` starts
# a comment.  The explicit-matrix gate is `u k lane_1 .. lane_k re im re im ...` with the 4^k entries column-major.
# First line: `qubits n`.
def dumps(n, gates) -> str:
    lines = [f"qubits {n}"]
    for g in gates:
        if g.name == "u":
            flat = np.reshape(g.matrix(), -1, order="F")
            nums = " ".join(f"{float(v.real)!r} {float(v.imag)!r}" for v in flat)
            lines.append(f"u {len(g.lanes)} {' '.join(map(str, g.lanes))} {nums}")
        else:
            lines.append(" ".join([g.name] + [str(l) for l in g.lanes] + [repr(p) for p in g.params]))
    return "\n".join(lines) + "\n"


def loads(text: str):
    """-> (n, [Gate]); raises ValueError with the line number on malformed input."""
    n, gates = None, []
    for ln, raw in enumerate(text.splitlines(), 1):
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        tok = line.split()
        try:
            name = tok[0].lower()
            if name == "qubits":
                n = int(tok[1])
                continue
            if n is None:
                raise ValueError("the first statement must be `qubits n`")
            if name == "u":
                k = int(tok[1])
                lanes = tuple(int(t) for t in tok[2:2 + k])
                vals = [float(t) for t in tok[2 + k:]]
                if len(vals) != 2 * 4 ** k:
                    raise ValueError(f"expected {2 * 4 ** k} numbers")
                m = np.reshape(np.array(vals[0::2]) + 1j * np.array(vals[1::2]), (2 ** k, 2 ** k), order="F")
                g = Gate("u", lanes, (), m)
            else:
                if name not in GATES:
                    raise ValueError(f"unknown gate '{name}'")
                nl = GATES[name][0]
                g = Gate(name, tuple(int(t) for t in tok[1:1 + nl]), tuple(float(t) for t in tok[1 + nl:]))
            if max(g.lanes) > n:
                raise ValueError(f"lane {max(g.lanes)} > qubits {n}")
            gates.append(g)
        except (IndexError, ValueError) as e:
            raise ValueError(f"gate list line {ln}: {e}") from None
    if n is None:
        raise ValueError("gate list: missing `qubits n`")
    return n, gates


def load(path):
    with open(path) as f:
        return loads(f.read())


def dump(path, n, gates):
    with open(path, "w") as f:
        f.write(dumps(n, gates))


def qft(n):
    """`Quac.Algorithms.QFT(n)` [ext]: H and controlled phases, then the bit-reversal SWAPs."""
    gates = []
    for j in range(1, n + 1):
        gates.append(Gate("h", (j,)))
        for k in range(j + 1, n + 1):
            gates.append(Gate("cphase", (k, j), (2 * math.pi / 2 ** (k - j + 1),)))
    for j in range(1, n // 2 + 1):
        gates.append(Gate("swap", (j, n + 1 - j)))
    return gates
