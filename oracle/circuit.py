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


# ---- deterministic planner (same rules as qrochet.jl_b200/csrc/tn.cu, compared bit-exactly by the tests) ----------
# The reference plans with `transform!(tn, ContractSimplification())` + `einexpr(tn; optimizer = HyPar(...))` +
# `findslices(SizeScorer(), path; size)` (examples/distributed.jl:29-46).  KaHyPar / EinExprs are un-vendored and
# randomised, so the rules are OURS and are stated here:
#   simplify    : while some pair of connected tensors contracts to a result no larger than its larger operand, contract
#                 the first such pair in (position, position) order of the live list and restart the scan
#                 (ContractSimplification: rank-1 boundary vectors and non-growing pairs are absorbed before planning);
#   greedy(α/2) : repeatedly contract the connected pair minimising 2 size(out) - α (size(a) + size(b)); ties -> smaller
#                 output -> earlier pair; disconnected remainder: outer product of the two smallest;
#   reconfigure : sub-tree reconfiguration (the local search): walk the tree from the root; at each internal node take
#                 the frontier of up to 8 sub-trees below it (repeatedly open the largest openable one), find by
#                 exhaustive dynamic programming over the 2^8 subsets the contraction order of those sub-trees with the
#                 fewest flops (ties: smaller largest intermediate), and adopt it when it is strictly better; rounds
#                 repeat until nothing improves (at most 32);
#   candidates  : {simplify on, off} x α in {2, 4}; each is sliced by findslices and the one with the fewest total
#                 flops (slices x per-slice + slice-invariant work) wins, the first on ties;
#   findslices  : while the largest intermediate exceeds max_elements, cut the index with the largest
#                 score = sum of the sizes of the nodes holding it; ties -> the index met first in post-order.
# All sizes and flop counts are exact integers (Python int here, unsigned __int128 in C++, both saturating at 2^100), so
# both implementations take identical decisions.
PLAN_K = 8
PLAN_ROUNDS = 32
PLAN_ALPHAS = (2, 4)
PLAN_CAP = 1 << 100        # sizes saturate here (unsigned __int128 head-room on the C++ side)
PLAN_CAP_TOTAL = 1 << 126


class _Net:
    def __init__(self, modes, extents):
        self.leaf_modes = [tuple(m) for m in modes]
        self.ext = extents
        self.total = {}
        for m in self.leaf_modes:
            for x in m:
                self.total[x] = self.total.get(x, 0) + 1

    def size(self, ms, cut=()):
        s = 1
        for x in ms:
            if x not in cut:
                s = min(s * self.ext[x], PLAN_CAP)
        return s

    def out_modes(self, A, B):
        """Modes of the node contracting A and B (without building it)."""
        ca, cb, tot = A["cnt"], B["cnt"], self.total
        out = [x for x in A["modes"] if ca.get(x, 0) + cb.get(x, 0) < tot[x]]
        out += [x for x in B["modes"] if x not in ca and cb[x] < tot[x]]
        return out

    def leaf(self, i):
        cnt = {}
        for x in self.leaf_modes[i]:
            cnt[x] = cnt.get(x, 0) + 1
        return dict(left=-1, right=-1, modes=self.leaf_modes[i], cnt=cnt)

    def join(self, nodes, a, b):
        """The node contracting nodes a and b: an index is summed once every leaf holding it is below the node."""
        A, B = nodes[a], nodes[b]
        cnt = dict(A["cnt"])
        for x, v in B["cnt"].items():
            cnt[x] = cnt.get(x, 0) + v
        out = [x for x in A["modes"] if cnt[x] < self.total[x]]
        out += [x for x in B["modes"] if x not in A["modes"] and cnt[x] < self.total[x]]
        return dict(left=a, right=b, modes=tuple(out), cnt=cnt)

    def flops(self, nodes, i, cut=()):
        """complex multiply-adds of node i = product of the extents of every index involved (EinExprs `flops`)."""
        n = nodes[i]
        inv = dict.fromkeys(nodes[n["left"]]["modes"] + nodes[n["right"]]["modes"])
        return self.size(inv, cut)


def _node_cost(objective, macs, elems):
    """0: complex multiply-adds; 1: the executor's time model on B200 in multiply-add units, max(macs, 10 x elements
    moved) (tn_plan.cu: node_cost)."""
    if objective == 0:
        return macs
    bw = (1 << 110) if elems > (1 << 110) // 10 else elems * 10
    return max(macs, bw)


def _connected(a, b):
    return any(x in b["modes"] for x in a["modes"])


def _simplify(net, nodes, live):
    found = True
    while found:
        found = False
        for xi in range(len(live)):
            for yi in range(xi + 1, len(live)):
                a, b = live[xi], live[yi]
                if not _connected(nodes[a], nodes[b]):
                    continue
                o = net.out_modes(nodes[a], nodes[b])
                if net.size(o) <= max(net.size(nodes[a]["modes"]), net.size(nodes[b]["modes"])):
                    nodes.append(net.join(nodes, a, b))
                    del live[yi]
                    del live[xi]
                    live.append(len(nodes) - 1)
                    found = True
                    break
            if found:
                break


def _greedy(net, nodes, live, alpha2):
    while len(live) > 1:
        best = None
        for xi in range(len(live)):
            a = nodes[live[xi]]
            sa = net.size(a["modes"])
            for yi in range(xi + 1, len(live)):
                b = nodes[live[yi]]
                if not _connected(a, b):
                    continue
                so = net.size(net.out_modes(a, b))
                cost = 2 * so - alpha2 * (sa + net.size(b["modes"]))
                if best is None or cost < best[0] or (cost == best[0] and so < best[1]):
                    best = (cost, so, xi, yi)
        if best is None:  # disconnected components: outer product of the two smallest
            idx = sorted(range(len(live)), key=lambda i: net.size(nodes[live[i]]["modes"]))
            xi, yi = min(idx[0], idx[1]), max(idx[0], idx[1])
        else:
            _, _, xi, yi = best
        nodes.append(net.join(nodes, live[xi], live[yi]))
        del live[yi]
        del live[xi]
        live.append(len(nodes) - 1)


def _reconfigure_at(net, nodes, i, objective=0):
    """One sub-tree reconfiguration at internal node i; True when the sub-tree was replaced by a cheaper one."""
    fr = [nodes[i]["left"], nodes[i]["right"]]
    while len(fr) < PLAN_K:
        pick = -1
        for pos, j in enumerate(fr):
            if nodes[j]["left"] >= 0 and (pick < 0 or net.size(nodes[j]["modes"]) > net.size(nodes[fr[pick]]["modes"])):
                pick = pos
        if pick < 0:
            break
        j = fr.pop(pick)
        fr += [nodes[j]["left"], nodes[j]["right"]]
    K = len(fr)
    if K < 3:
        return False
    frontier = set(fr)
    old_fl, old_mx = 0, 0
    stack = [i]
    while stack:
        j = stack.pop()
        if j in frontier:
            continue
        old_fl += _node_cost(objective, net.flops(nodes, j),
                             net.size(nodes[nodes[j]["left"]]["modes"]) + net.size(nodes[nodes[j]["right"]]["modes"])
                             + net.size(nodes[j]["modes"]))
        old_mx = max(old_mx, net.size(nodes[j]["modes"]))
        stack += [nodes[j]["left"], nodes[j]["right"]]
    full = (1 << K) - 1
    # relevant modes: the outer indices of the frontier sub-trees (everything else is summed inside one of them)
    rel = list(dict.fromkeys(x for j in fr for x in nodes[j]["modes"]))
    cnt, modes, size = {}, {}, {}
    for m in range(1, full + 1):
        low = (m & -m).bit_length() - 1
        rest = m & (m - 1)
        c = nodes[fr[low]]["cnt"]
        cnt[m] = [c.get(x, 0) + (cnt[rest][r] if rest else 0) for r, x in enumerate(rel)]
        modes[m] = frozenset(x for r, x in enumerate(rel) if 0 < cnt[m][r] < net.total[x])
        size[m] = net.size(modes[m])
    best = {1 << j: (0, 0, None) for j in range(K)}
    for m in range(1, full + 1):
        if m & (m - 1) == 0:
            continue
        low = m & -m
        choice = None
        sub = (m - 1) & m
        while sub:
            if sub & low:
                o = m ^ sub
                fl = best[sub][0] + best[o][0] + _node_cost(objective, net.size(modes[sub] | modes[o]),
                                                            size[sub] + size[o] + size[m])
                mx = max(best[sub][1], best[o][1], size[m])
                if choice is None or (fl, mx) < (choice[0], choice[1]):
                    choice = (fl, mx, (sub, o))
            sub = (sub - 1) & m
        best[m] = choice
    if (best[full][0], best[full][1]) >= (old_fl, old_mx):
        return False

    def build(m):
        if m & (m - 1) == 0:
            return fr[m.bit_length() - 1]
        sub, o = best[m][2]
        ia = build(sub)
        ib = build(o)
        nd = net.join(nodes, ia, ib)
        if m == full:
            nodes[i] = nd
            return i
        nodes.append(nd)
        return len(nodes) - 1

    build(full)
    return True


def _reconfigure(net, nodes, root, objective=0):
    for _ in range(PLAN_ROUNDS):
        improved = False
        stack = [root]
        while stack:
            i = stack.pop()
            if nodes[i]["left"] < 0:
                continue
            if _reconfigure_at(net, nodes, i, objective):
                improved = True
            stack += [nodes[i]["right"], nodes[i]["left"]]  # left sub-tree first
        if not improved:
            break


def _compact(net, nodes, root):
    """Post-order renumbering of the reachable tree: leaves keep their ids, step s creates node nleaves + s."""
    nleaves = len(net.leaf_modes)
    path, newid = [], {}
    stack = [(root, False)]
    while stack:
        i, done = stack.pop()
        if nodes[i]["left"] < 0:
            newid[i] = i
        elif done:
            path.append((newid[nodes[i]["left"]], newid[nodes[i]["right"]]))
            newid[i] = nleaves + len(path) - 1
        else:
            stack += [(i, True), (nodes[i]["right"], False), (nodes[i]["left"], False)]
    out = [net.leaf(i) for i in range(nleaves)]
    for a, b in path:
        out.append(net.join(out, a, b))
    return out, path


def _findslices(net, nodes, max_elements):
    nleaves = len(net.leaf_modes)
    cut = []
    if max_elements <= 0:
        return cut
    while True:
        mx = max((net.size(nodes[i]["modes"], cut) for i in range(nleaves, len(nodes))), default=0)
        if mx <= max_elements:
            break
        score, seen = {}, []
        for i in range(len(nodes)):  # compacted trees are stored in post-order
            s = net.size(nodes[i]["modes"], cut)
            for x in nodes[i]["modes"]:
                if x in cut or net.ext[x] <= 1:
                    continue
                if x not in score:
                    seen.append(x)
                    score[x] = 0
                score[x] += s
        if not seen:
            break
        pick = seen[0]
        for x in seen:
            if score[x] > score[pick]:
                pick = x
        cut.append(pick)
    return cut


def _sliced_cost(net, nodes, cut, objective=0):
    """(complex MACs per slice of the nodes that depend on a cut index, MACs of the slice-invariant nodes, #slices,
    objective cost per slice, objective cost of the invariant nodes)."""
    nleaves = len(net.leaf_modes)
    inv = [not any(x in cut for x in net.leaf_modes[i]) for i in range(nleaves)]
    per_slice = once = obj_ps = obj_once = 0
    for i in range(nleaves, len(nodes)):
        inv.append(inv[nodes[i]["left"]] and inv[nodes[i]["right"]])
        f = net.flops(nodes, i, cut)
        c = _node_cost(objective, f, net.size(nodes[nodes[i]["left"]]["modes"], cut)
                       + net.size(nodes[nodes[i]["right"]]["modes"], cut) + net.size(nodes[i]["modes"], cut))
        if inv[i]:
            once += f
            obj_once += c
        else:
            per_slice += f
            obj_ps += c
    nsl = 1
    for x in cut:
        nsl = min(nsl * net.ext[x], 1 << 62)
    return per_slice, once, nsl, obj_ps, obj_once


def plan(modes, extents, max_elements, optimizer=1):
    """modes: list of tuples; extents: dict mode -> size.  optimizer 0: the round-1 rule (one greedy tree, α = 1, no
    simplification, no local search); 1: the full planner above minimising flops; 2: the same minimising the executor's
    time model (_node_cost).  Returns dict(path, sliced, nodes, macs_per_slice,
    macs_invariant, nslices)."""
    net = _Net(modes, extents)
    nleaves = len(net.leaf_modes)
    best = None
    for simp in ((1, 0) if optimizer else (0,)):
        for alpha2 in (PLAN_ALPHAS if optimizer else (2,)):
            nodes = [net.leaf(i) for i in range(nleaves)]
            live = list(range(nleaves))
            if simp:
                _simplify(net, nodes, live)
            _greedy(net, nodes, live, alpha2)
            root = live[0]
            objective = 1 if optimizer >= 2 else 0
            if optimizer:
                _reconfigure(net, nodes, root, objective)
            tree, path = _compact(net, nodes, root)
            cut = _findslices(net, tree, max_elements)
            per_slice, once, nsl, obj_ps, obj_once = _sliced_cost(net, tree, cut, objective)
            total = min(obj_ps * nsl + obj_once, PLAN_CAP_TOTAL)
            if best is None or total < best[0]:
                best = (total, tree, path, cut, per_slice, once, nsl)
    _, tree, path, cut, per_slice, once, nsl = best
    return dict(path=path, sliced=cut, nodes=tree, macs_per_slice=per_slice, macs_invariant=once, nslices=nsl)


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
