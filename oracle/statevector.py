"""ORACLE (test infrastructure, NOT product code) -- dense state-vector known answers
for the paths the reference leaves untested (evolve!/expect/overlap, `Chain_test.jl:396`;
sliced amplitude contraction, `examples/distributed.jl`).  Qubit/site 1 is the fastest
index of the vector (column-major over sites), matching `Chain.to_dense`.
"""
from __future__ import annotations

import numpy as np


def apply_gate(psi, matrix, sites, n, p=2):
    """psi <- (matrix on `sites`, 1-based, first listed site = fastest bit of the matrix index)."""
    k = len(sites)
    t = np.reshape(psi, (p,) * n, order="F")
    g = np.reshape(np.asarray(matrix), (p,) * (2 * k), order="F")  # (o_1..o_k, i_1..i_k)
    axes = [s - 1 for s in sites]
    t = np.tensordot(g, t, axes=(list(range(k, 2 * k)), axes))  # (o_1..o_k, rest...)
    rest = [a for a in range(n) if a not in axes]
    # put the output axes back to their site positions
    order = [0] * n
    for pos, a in enumerate(axes):
        order[a] = pos
    for pos, a in enumerate(rest):
        order[a] = k + pos
    t = np.transpose(t, order)
    return np.reshape(t, -1, order="F")


def expect(psi, matrix, sites, n, p=2):
    """<psi| O |psi>, un-normalised (Chain.jl:724-735 semantics)."""
    return np.vdot(psi, apply_gate(psi, matrix, sites, n, p))


def zero_state(n, p=2):
    psi = np.zeros(p ** n, dtype=np.complex128)
    psi[0] = 1.0
    return psi


def fsim(theta, phi):
    """FSim(θ,φ) of Yao.EasyBuild [ext]: [[1,0,0,0],[0,cosθ,-i sinθ,0],[0,-i sinθ,cosθ,0],[0,0,0,e^{-iφ}]]."""
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1, 0, 0, 0], [0, c, -1j * s, 0], [0, -1j * s, c, 0], [0, 0, 0, np.exp(-1j * phi)]],
                    dtype=np.complex128)
