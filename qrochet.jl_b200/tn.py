"""Sliced contraction of a circuit tensor network (examples/distributed.jl of the reference) on B200.

Host side mirrors the reference's glue: `Quantum(circuit)` (ext/QrochetYaoExt.jl:15-48: one rank-2k tensor per
gate with array dims labelled [in.., out..]), `zeros(Product, n)` boundaries (src/Ansatz/Product.jl:44-46),
`merge` into the closed network <0|U|0> (examples/distributed.jl:25-28).  Planning (path + slicing) and the
per-slice tree replay run inside libqrochet_b200.so; ranks take slices s = rank, rank+W, ... and one NCCL sum
finishes the job."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import check, lib
from .device import Context


def fsim(theta, phi):
    """FSim(θ,φ) gate of Yao.EasyBuild [ext] (examples/distributed.jl:20)."""
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [0, 0, 0, np.exp(-1j * phi)]],
                    dtype=np.complex128)


def random_fsim_circuit(n, depth, seed=3000):
    """`depth` layers of FSim(2πu, 2πu') on a random perfect matching (examples/distributed.jl:14-23), seeded
    per layer with default_rng(seed + layer).  Returns [(q1, q2, matrix)], qubits 0-based."""
    gates = []
    for layer in range(depth):
        rng = np.random.default_rng(seed + layer)
        perm = rng.permutation(n)
        for a in range(0, n - 1, 2):
            i, j = int(perm[a]), int(perm[a + 1])
            theta, phi = 2 * np.pi * rng.random(), 2 * np.pi * rng.random()
            gates.append((i, j, fsim(theta, phi)))
    return gates


def circuit_network(n, gates):
    """`Quantum(circuit)` of the reference's circuit front-ends (ext/QrochetYaoExt.jl:15-48, ext/QrochetQuacExt.jl:17-48):
    one rank-2k tensor per k-qubit gate, array = reshape(matrix, 2,...,2) with dims labelled [in_1..in_k, out_1..out_k]
    (as the reference does, :30-36), a SWAP only exchanges the wires.  `gates` = [(qubits, matrix)] with 0-based
    qubit tuples and matrix = None / "swap" for a SWAP.  Returns (arrays, modes, inputs, outputs): inputs[q] /
    outputs[q] are the open mode labels of wire q (Site(q; dual=true) / Site(q))."""
    nxt = [0]

    def fresh():
        nxt[0] += 1
        return nxt[0] - 1

    wire = [[fresh()] for _ in range(n)]
    arrays, modes = [], []
    for qubits, mat in gates:
        qubits = tuple(int(q) for q in qubits)
        if mat is None or (isinstance(mat, str) and mat == "swap"):
            a, b = qubits
            wire[a], wire[b] = wire[b], wire[a]
            continue
        k = len(qubits)
        mat = np.asarray(mat, dtype=np.complex128)
        assert mat.shape == (2 ** k, 2 ** k)
        froms, tos = [], []
        for q in qubits:
            froms.append(wire[q][-1])
            wire[q].append(fresh())
            tos.append(wire[q][-1])
        arrays.append(np.reshape(mat, (2,) * (2 * k), order="F"))
        modes.append(tuple(froms + tos))
    return arrays, modes, [w[0] for w in wire], [w[-1] for w in wire]


def amplitude_network(n, gates, ket=None, bra=None):
    """Leaves (arrays, modes) of the closed network <bra|U|ket>; `ket` / `bra` are product states given as lists of
    n local vectors (default `zeros(Product, n)`, i.e. <0..0|U|0..0> as in examples/distributed.jl:26-28; the bra
    enters conjugated, as `Quantum(ψ)'` does)."""
    nxt = [0]

    def fresh():
        nxt[0] += 1
        return nxt[0] - 1

    wire = [fresh() for _ in range(n)]
    first = list(wire)
    arrays, modes = [], []
    for (i, j, mat) in gates:
        arrays.append(np.reshape(np.asarray(mat, dtype=np.complex128), (2, 2, 2, 2), order="F"))
        fi, fj = wire[i], wire[j]
        wire[i], wire[j] = fresh(), fresh()
        modes.append((fi, fj, wire[i], wire[j]))
    zero = np.array([1.0, 0.0], dtype=np.complex128)
    for q in range(n):
        arrays.append(zero.copy() if ket is None else np.asarray(ket[q], dtype=np.complex128))
        modes.append((first[q],))
    for q in range(n):
        arrays.append(zero.copy() if bra is None else np.conj(np.asarray(bra[q], dtype=np.complex128)))
        modes.append((wire[q],))
    return arrays, modes


class SlicedContraction:
    """Plan once, replay the tree per slice (qb200_tn_plan_opt / qb200_tn_contract_sliced).  optimizer: 1 = full planner
    (simplification + multi-start greedy + sub-tree reconfiguration), 0 = the round-1 single greedy tree."""

    def __init__(self, ctx: Context | None, arrays, modes, max_elements: int, optimizer: int = 1):
        self.ctx = ctx
        self.modes = [tuple(int(x) for x in m) for m in modes]
        shapes = [tuple(np.shape(a)) for a in arrays]
        ranks = capi.i32arr(len(m) for m in self.modes)
        flat_modes = capi.i32arr(x for m in self.modes for x in m)
        flat_ext = capi.i64arr(e for s in shapes for e in s)
        h = C.c_void_p()
        ch = ctx.h if ctx is not None else None
        check(ch, lib.qb200_tn_plan_opt(ch, len(arrays), ranks, flat_modes, flat_ext, int(max_elements), int(optimizer),
                                        C.byref(h)))
        self.h = h
        self.leaves = None
        if ctx is not None:
            self.leaves = [ctx.array(np.asarray(a, dtype=np.complex128)) for a in arrays]
            self._leaf_ptrs = (C.c_void_p * len(arrays))(*[l.h for l in self.leaves])

    def __del__(self):
        try:
            if self.h:
                lib.qb200_tn_plan_free(self.ctx.h if self.ctx is not None else None, self.h)
        except Exception:
            pass
        self.h = None

    @property
    def nslices(self) -> int:
        return int(lib.qb200_tn_plan_nslices(self.h))

    @property
    def sliced_modes(self):
        n = lib.qb200_tn_plan_sliced_modes(self.h, None)
        buf = (C.c_int32 * max(n, 1))()
        lib.qb200_tn_plan_sliced_modes(self.h, buf)
        return [int(buf[i]) for i in range(n)]

    @property
    def path(self):
        n = lib.qb200_tn_plan_path(self.h, None)
        buf = (C.c_int32 * max(2 * n, 1))()
        lib.qb200_tn_plan_path(self.h, buf)
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)]

    @property
    def flops_per_slice(self) -> float:
        return float(lib.qb200_tn_plan_flops_per_slice(self.h))

    @property
    def flops_invariant(self) -> float:
        """8 x complex MACs of the slice-invariant tree nodes (executed once per contract call)."""
        return float(lib.qb200_tn_plan_flops_invariant(self.h))

    @property
    def max_intermediate(self) -> int:
        return int(lib.qb200_tn_plan_max_intermediate(self.h))

    def contract(self, first_slice=0, stride=1) -> complex:
        """Sum of the slices first_slice, first_slice + stride, ... (this rank's share)."""
        acc = (C.c_double * 2)(0.0, 0.0)
        check(self.ctx.h, lib.qb200_tn_contract_sliced(self.ctx.h, self.h, self._leaf_ptrs, int(first_slice),
                                                       int(stride), acc))
        return complex(acc[0], acc[1])
