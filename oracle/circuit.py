"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the sliced circuit-amplitude path of
`/root/reference/examples/distributed.jl` (with the two defects of that example corrected):

  * circuit -> tensor network as `ext/QrochetYaoExt.jl:15-48` does it (one rank-2k tensor per k-qubit gate,
    array dims labelled [in_1..in_k, out_1..out_k]), boundary <0|...|0> as `zeros(Product, n)`
    (`src/Ansatz/Product.jl:44-46`), merged as `examples/distributed.jl:25-28`;
  * a deterministic greedy contraction path and a deterministic `findslices(SizeScorer)`:
    EinExprs/KaHyPar are un-vendored, un-pinned and randomised [ext], so the choice rule is OURS and is stated
    here; libqrochet_b200's C++ planner implements the same rule and the tests require the same path and the
    same cut indices (bit-exact);
  * slices enumerated with the first cut index fastest (`examples/distributed.jl:47,69`), slice s on rank
    s mod W, partial sums added (`:79-101`).

PARITY UNPINNED by the reference (no test exercises this path): known answers are the dense state-vector
amplitude (`oracle/statevector.py`) and "sum over all slices == unsliced contraction".
"""
from __future__ import annotations

import numpy as np

from . import statevector as sv


def random_fsim_circuit(n, depth, seed=3000):
    """`depth` layers, each a random perfect matching (`default_rng(seed+layer).permutation(n)`) of
    FSim(2πu, 2πu') gates (examples/distributed.jl:14-23).  Returns [(q1, q2, 4x4 matrix)], qubits 0-based."""
    gates = []
    for layer in range(depth):
        rng = np.random.default_rng(seed + layer)
        perm = rng.permutation(n)
        for a in range(0, n - 1, 2):
            i, j = int(perm[a]), int(perm[a + 1])
            theta, phi = 2 * np.pi * rng.random(), 2 * np.pi * rng.random()
            gates.append((i, j, sv.fsim(theta, phi)))
    return gates


def amplitude_network(n, gates, ket=None, bra=None):
    """Closed network <bra| U |ket> (product states as lists of local vectors; default |0..0>) as (arrays, modes):
    modes are integer labels."""
    counter = [0]

    def fresh():
        counter[0] += 1
        return counter[0] - 1

    wire = [fresh() for _ in range(n)]
    first = list(wire)
    arrays, modes = [], []
    for (i, j, mat) in gates:
        arr = np.reshape(np.asarray(mat, dtype=np.complex128), (2, 2, 2, 2), order="F")
        fi, fj = wire[i], wire[j]
        ti, tj = fresh(), fresh()
        wire[i], wire[j] = ti, tj
        arrays.append(arr)
        modes.append((fi, fj, ti, tj))  # [in_1, in_2, out_1, out_2] (QrochetYaoExt.jl:30-36)
    zero = np.array([1.0, 0.0], dtype=np.complex128)
    for q in range(n):
        arrays.append(zero.copy() if ket is None else np.asarray(ket[q], dtype=np.complex128))
        modes.append((first[q],))
    for q in range(n):
        arrays.append(zero.copy() if bra is None else np.conj(np.asarray(bra[q], dtype=np.complex128)))
        modes.append((wire[q],))
    return arrays, modes


def product_vector(vectors):
    """Dense vector of a product state, site 1 the fastest index."""
    psi = np.ones(1, dtype=np.complex128)
    for v in vectors:
        psi = np.kron(np.asarray(v, dtype=np.complex128), psi)
    return psi


def statevector_amplitude(n, gates, ket=None, bra=None):
    psi = sv.zero_state(n) if ket is None else product_vector(ket)
    for (i, j, mat) in gates:
        # the Yao extension labels array dims 1..k as the incoming wires (= transposed gate); FSim is symmetric
        psi = sv.apply_gate(psi, np.asarray(mat).T, [i + 1, j + 1], n)
    return psi[0] if bra is None else np.vdot(product_vector(bra), psi)


def random_product_state(n, seed):
    """n normalised random local vectors (FSim circuits act trivially on |0..0>, so <0|U|0> = 1 is a weak known
    answer: the tests use random product states on both sides)."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 2)) + 1j * rng.standard_normal((n, 2))
    return list(v / np.linalg.norm(v, axis=1, keepdims=True))


# ---- deterministic planner (same rule as qrochet.jl_b200/csrc/tn.cu) -----------------------------------
def plan(modes, extents, max_elements):
    """modes: list of tuples; extents: dict mode -> size.  Returns dict(path, sliced, nodes)."""
    nodes = [dict(modes=tuple(m), left=-1, right=-1) for m in modes]
    count = {}
    for m in modes:
        for x in m:
            count[x] = count.get(x, 0) + 1

    def size(ms):
        s = 1
        for x in ms:
            s *= extents[x]
        return s

    def out_modes(a, b):
        out = []
        for x in a:
            if count[x] - 1 - (1 if x in b else 0) > 0:
                out.append(x)
        for x in b:
            if x not in a and count[x] - 1 > 0:
                out.append(x)
        return tuple(out)

    live = list(range(len(nodes)))
    path = []
    while len(live) > 1:
        best = None
        for xi in range(len(live)):
            a = nodes[live[xi]]["modes"]
            sa = set(a)
            for yi in range(xi + 1, len(live)):
                b = nodes[live[yi]]["modes"]
                if not sa.intersection(b):
                    continue
                o = out_modes(a, b)
                so = size(o)
                cost = so - size(a) - size(b)
                if best is None or cost < best[0] or (cost == best[0] and so < best[1]):
                    best = (cost, so, xi, yi, o)
        if best is None:
            idx = sorted(range(len(live)), key=lambda i: size(nodes[live[i]]["modes"]))
            xi, yi = min(idx[0], idx[1]), max(idx[0], idx[1])
            o = out_modes(nodes[live[xi]]["modes"], nodes[live[yi]]["modes"])
        else:
            _, _, xi, yi, o = best
        ia, ib = live[xi], live[yi]
        for x in nodes[ia]["modes"]:
            count[x] -= 1
        for x in nodes[ib]["modes"]:
            count[x] -= 1
        for x in o:
            count[x] += 1
        nodes.append(dict(modes=o, left=ia, right=ib))
        path.append((ia, ib))
        del live[yi]
        del live[xi]
        live.append(len(nodes) - 1)

    order = []

    def post(i):
        if nodes[i]["left"] >= 0:
            post(nodes[i]["left"])
            post(nodes[i]["right"])
        order.append(i)

    post(len(nodes) - 1)
    cut = []

    def nsize(ms):
        s = 1
        for x in ms:
            if x not in cut:
                s *= extents[x]
        return s

    nleaves = len(modes)
    if max_elements > 0:
        while True:
            mx = max((nsize(nodes[i]["modes"]) for i in range(nleaves, len(nodes))), default=0)
            if mx <= max_elements:
                break
            score, seen = {}, []
            for i in order:
                s = float(nsize(nodes[i]["modes"]))
                for x in nodes[i]["modes"]:
                    if x in cut or extents[x] <= 1:
                        continue
                    if x not in score:
                        seen.append(x)
                        score[x] = 0.0
                    score[x] += s
            if not seen:
                break
            pick = seen[0]
            for x in seen:
                if score[x] > score[pick]:
                    pick = x
            cut.append(pick)
    return dict(path=path, sliced=cut, nodes=nodes)


def contract_sliced(arrays, modes, pl, first_slice=0, stride=1):
    """Sum over slices first_slice, first_slice+stride, ... of the fixed tree (first cut index fastest)."""
    extents = {}
    for a, m in zip(arrays, modes):
        for x, e in zip(m, a.shape):
            extents[x] = e
    cut = pl["sliced"]
    nsl = int(np.prod([extents[x] for x in cut], dtype=np.int64)) if cut else 1
    acc = 0.0 + 0.0j
    nleaves = len(arrays)
    for s in range(first_slice, nsl, stride):
        rem, val = s, {}
        for x in cut:
            val[x] = rem % extents[x]
            rem //= extents[x]
        bufs = {}
        for t, (a, m) in enumerate(zip(arrays, modes)):
            mm = list(m)
            for x in cut:
                if x in mm:
                    ax = mm.index(x)
                    a = np.take(a, val[x], axis=ax)
                    del mm[ax]
            bufs[t] = (a, tuple(mm))
        for step, (ia, ib) in enumerate(pl["path"]):
            node = pl["nodes"][nleaves + step]
            om = tuple(x for x in node["modes"] if x not in cut)
            (a, ma), (b, mb) = bufs.pop(ia), bufs.pop(ib)
            letters = {}
            for x in ma + mb:
                letters.setdefault(x, len(letters))
            res = np.einsum(a, [letters[x] for x in ma], b, [letters[x] for x in mb], [letters[x] for x in om])
            bufs[nleaves + step] = (res, om)
        (res, om), = bufs.values()
        acc += complex(res)
    return acc, nsl
