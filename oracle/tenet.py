"""ORACLE (test infrastructure, NOT product code) -- labelled tensors and a tensor
network with hyper-index semantics, restating the part of Tenet.jl v0.6 that
Qrochet.jl calls.

Tenet itself is an un-vendored, un-pinned dependency of the reference
(`/root/reference/Project.toml:32`, compat "0.6"; Manifest git-ignored), so its
source is NOT under /root/reference.  The semantics restated here are the ones
listed in SURVEY.md Appendix C and they are anchored on the reference's own
call sites (cited per function) and on the reference's property tests
(`test/Ansatz/Chain_test.jl`), which `tests/test_oracle_properties.py` restates.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.

Numerics: NumPy/SciPy on OpenBLAS -> zgemm / LAPACK zgesdd / zgeqrf, the same
LAPACK drivers Julia's `LinearAlgebra.svd` / `qr` dispatch to.
"""
from __future__ import annotations

import itertools
from typing import Iterable, Sequence

import numpy as np
import scipy.linalg as sla

_counter = itertools.count()


def nextindex() -> str:
    """Fresh index label; restates `Qrochet.nextindex` (src/Utils.jl:3-7)."""
    return f"i{next(_counter)}"


def gensym(tag: str = "tmp") -> str:
    return f"##{tag}#{next(_counter)}"


class Tensor:
    """`Tenet.Tensor(data, inds)`: dense array + one label per axis."""

    __slots__ = ("data", "inds")

    def __init__(self, data, inds: Sequence[str]):
        data = np.asarray(data)
        inds = tuple(inds)
        assert data.ndim == len(inds), (data.shape, inds)
        assert len(set(inds)) == len(inds), f"repeated index in {inds}"
        self.data = data
        self.inds = inds

    def size(self, ind: str) -> int:
        return self.data.shape[self.inds.index(ind)]

    @property
    def shape(self):
        return self.data.shape

    def conj(self) -> "Tensor":
        return Tensor(np.conj(self.data), self.inds)

    def copy(self) -> "Tensor":
        return Tensor(self.data.copy(), self.inds)

    def replace(self, mapping: dict) -> "Tensor":
        return Tensor(self.data, tuple(mapping.get(i, i) for i in self.inds))

    def permute(self, inds: Sequence[str]) -> "Tensor":
        inds = tuple(inds)
        return Tensor(np.transpose(self.data, [self.inds.index(i) for i in inds]), inds)

    def __repr__(self):
        return f"Tensor{self.inds}{self.data.shape}"


def _letters(all_inds):
    table = {}
    for i in all_inds:
        if i not in table:
            table[i] = len(table)
    return table


def contract(a: Tensor, b: Tensor, dims=None) -> Tensor:
    """`Tenet.contract(a, b; dims)`.

    Sums `dims ∩ inds(a) ∩ inds(b)` (default: every shared index); `dims=()`
    keeps shared indices (element-wise along them).  Output index order is
    `(inds(a) ∪ inds(b)) \\ dims`.  Call sites: Chain.jl:322,325,435,453,484,
    491,710,713.
    """
    shared = [i for i in a.inds if i in b.inds]
    dims = shared if dims is None else [i for i in dims if i in shared]
    out = [i for i in a.inds if i not in dims] + [i for i in b.inds if i not in a.inds]
    t = _letters(a.inds + b.inds)
    # optimize=True: a pairwise contraction goes through tensordot (BLAS zgemm, what OMEinsum's TTGT path calls)
    res = np.einsum(a.data, [t[i] for i in a.inds], b.data, [t[i] for i in b.inds], [t[i] for i in out], optimize=True)
    return Tensor(res, out)


def contract_many(tensors: Sequence[Tensor], summed: Iterable[str]) -> Tensor:
    """Contract several tensors at once summing exactly `summed` (hyper-index aware)."""
    summed = set(summed)
    all_inds = [i for t in tensors for i in t.inds]
    t = _letters(all_inds)
    out = [i for i in t if i not in summed]
    args = []
    for x in tensors:
        args += [x.data, [t[i] for i in x.inds]]
    res = np.einsum(*args, [t[i] for i in out], optimize="greedy" if len(tensors) > 1 else False)  # pairwise steps -> BLAS
    return Tensor(res, out)


def _matricize(t: Tensor, left_inds, right_inds):
    left_inds, right_inds = list(left_inds), list(right_inds)
    assert set(left_inds) | set(right_inds) == set(t.inds) and not set(left_inds) & set(right_inds)
    p = t.permute(left_inds + right_inds).data
    lshape = p.shape[: len(left_inds)]
    rshape = p.shape[len(left_inds):]
    # Julia reshape is column-major: first index fastest
    mat = np.reshape(p, (int(np.prod(lshape, dtype=np.int64)), int(np.prod(rshape, dtype=np.int64))), order="F")
    return mat, lshape, rshape


def svd(t: Tensor, left_inds, right_inds, virtualind: str):
    """`LinearAlgebra.svd(::Tensor; left_inds, right_inds, virtualind)` (call sites
    Chain.jl:365,645,705): thin SVD by LAPACK gesdd of the (left|right)
    matricisation; returns U[left..., v], s[v] (descending), Vt[right..., v] = conj(V)."""
    left_inds = list(left_inds)
    right_inds = list(right_inds) if right_inds else [i for i in t.inds if i not in left_inds]
    if not left_inds:
        left_inds = [i for i in t.inds if i not in right_inds]
    mat, lshape, rshape = _matricize(t, left_inds, right_inds)
    u, s, vh = sla.svd(mat, full_matrices=False, lapack_driver="gesdd")
    k = s.shape[0]
    U = Tensor(np.reshape(u, lshape + (k,), order="F"), left_inds + [virtualind])
    S = Tensor(s, [virtualind])
    # conj(V)[r, v] = Vh[v, r]
    Vt = Tensor(np.reshape(vh.T, rshape + (k,), order="F"), right_inds + [virtualind])
    return U, S, Vt


def qr(t: Tensor, left_inds, right_inds, virtualind: str):
    """`LinearAlgebra.qr(::Tensor; ...)` (call site Chain.jl:367): thin Householder QR;
    Q[left..., v], R[v, right...]."""
    left_inds = list(left_inds)
    right_inds = list(right_inds) if right_inds else [i for i in t.inds if i not in left_inds]
    if not left_inds:
        left_inds = [i for i in t.inds if i not in right_inds]
    mat, lshape, rshape = _matricize(t, left_inds, right_inds)
    q, r = sla.qr(mat, mode="economic")
    k = q.shape[1]
    Q = Tensor(np.reshape(q, lshape + (k,), order="F"), left_inds + [virtualind])
    R = Tensor(np.reshape(r, (k,) + rshape, order="F"), [virtualind] + right_inds)
    return Q, R


class TensorNetwork:
    """`Tenet.TensorNetwork`: a bag of tensors; equal labels are connected.
    open index = in one tensor, inner = two, hyper = three or more."""

    def __init__(self, tensors: Iterable[Tensor] = ()):
        self.tensors: list[Tensor] = list(tensors)

    def copy(self) -> "TensorNetwork":
        return TensorNetwork(self.tensors)  # shallow: shares arrays (Quantum.jl:87)

    def deepcopy(self) -> "TensorNetwork":
        return TensorNetwork([t.copy() for t in self.tensors])

    def inds(self, set: str = "all"):
        count: dict = {}
        for t in self.tensors:
            for i in t.inds:
                count[i] = count.get(i, 0) + 1
        if set == "all":
            return list(count)
        if set == "open":
            return [i for i, c in count.items() if c == 1]
        if set == "inner":
            return [i for i, c in count.items() if c >= 2]
        if set == "hyper":
            return [i for i, c in count.items() if c >= 3]
        raise ValueError(set)

    def size(self, ind: str) -> int:
        for t in self.tensors:
            if ind in t.inds:
                return t.size(ind)
        raise KeyError(ind)

    def intersecting(self, ind: str):
        return [t for t in self.tensors if ind in t.inds]

    def select(self, inds: Iterable[str]) -> Tensor:
        """`tn[inds...]`: the tensor whose index *set* equals `inds` (Ansatz.jl:88, Chain.jl:401)."""
        want = frozenset(inds)
        hits = [t for t in self.tensors if frozenset(t.inds) == want]
        if len(hits) != 1:
            raise KeyError(f"{len(hits)} tensors with inds {sorted(want)}")
        return hits[0]

    def push(self, t: Tensor):
        self.tensors.append(t)

    def delete(self, t: Tensor):
        for k, x in enumerate(self.tensors):
            if x is t:
                del self.tensors[k]
                return
        raise KeyError("tensor not in network")

    pop = delete

    def replace_tensor(self, old: Tensor, new: Tensor):
        for k, x in enumerate(self.tensors):
            if x is old:
                self.tensors[k] = new
                return
        raise KeyError("tensor not in network")

    def replace_inds(self, mapping: dict):
        self.tensors = [t.replace(mapping) if any(i in mapping for i in t.inds) else t for t in self.tensors]

    def merge(self, other: "TensorNetwork") -> "TensorNetwork":
        return TensorNetwork(self.tensors + other.tensors)

    def merge_(self, other: "TensorNetwork"):
        self.tensors += other.tensors
        return self

    def conj(self) -> "TensorNetwork":
        return TensorNetwork([t.conj() for t in self.tensors])

    def contract_index(self, inds):
        """`contract!(tn, i)` (call sites Chain.jl:372,602,616,636,682): remove every tensor
        touching any of `inds`, contract them together summing those indices, push the result."""
        inds = [inds] if isinstance(inds, str) else list(inds)
        hit = [t for t in self.tensors if any(i in t.inds for i in inds)]
        if not hit:
            return self
        rest = [t for t in self.tensors if not any(t is h for h in hit)]
        res = hit[0] if len(hit) == 1 and not inds else contract_many(hit, inds)
        self.tensors = rest + [res]
        return res

    def slice_(self, ind: str, keep: Sequence[int]):
        """`slice!(tn, ind, range)` (Chain.jl:419): restrict `ind` to the (0-based) positions
        `keep` on every tensor holding it; the index is kept."""
        keep = list(keep)
        out = []
        for t in self.tensors:
            if ind in t.inds:
                out.append(Tensor(np.take(t.data, keep, axis=t.inds.index(ind)), t.inds))
            else:
                out.append(t)
        self.tensors = out

    def view(self, fixed: dict) -> "TensorNetwork":
        """`view(tn, ind => value ...)` with integer values: the index is dropped
        (examples/distributed.jl:72,82)."""
        out = []
        for t in self.tensors:
            data, inds = t.data, list(t.inds)
            for i, v in fixed.items():
                if i in inds:
                    ax = inds.index(i)
                    data = np.take(data, v, axis=ax)
                    del inds[ax]
            out.append(Tensor(data, inds))
        return TensorNetwork(out)

    def svd_(self, left_inds, right_inds, virtualind):
        """`svd!(tn; left_inds, right_inds, virtualind)`: factorise `tn[left ∪ right]` in place;
        pushes U, s, Vt (s sits on the now hyper index)."""
        t = self.select(list(left_inds) + list(right_inds))
        U, S, Vt = svd(t, left_inds, right_inds, virtualind)
        self.delete(t)
        self.tensors += [U, S, Vt]
        return U, S, Vt

    def qr_(self, left_inds, right_inds, virtualind):
        t = self.select(list(left_inds) + list(right_inds))
        Q, R = qr(t, left_inds, right_inds, virtualind)
        self.delete(t)
        self.tensors += [Q, R]
        return Q, R

    def contract(self) -> Tensor:
        """`contract(tn)`: full contraction; an index is summed when every tensor holding it has
        been merged (hyper-index aware), open indices survive.  Greedy pairwise order."""
        opened = set(self.inds("open"))
        work = list(self.tensors)
        if not work:
            raise ValueError("empty network")
        # how many live tensors hold each index (an index survives a pairwise step while someone else still holds it)
        holders = {}
        for t in work:
            for i in set(t.inds):
                holders[i] = holders.get(i, 0) + 1
        while len(work) > 1:
            # cheapest-result-first greedy choice among connected pairs
            best = None
            sets = [set(t.inds) for t in work]
            for a, b in itertools.combinations(range(len(work)), 2):
                if not sets[a] & sets[b]:
                    continue
                ia, ib = work[a].inds, work[b].inds
                keep = [i for i in dict.fromkeys(ia + ib)
                        if i in opened or holders[i] - (i in sets[a]) - (i in sets[b]) > 0]
                size = 1
                for i in keep:
                    size *= work[a].size(i) if i in sets[a] else work[b].size(i)
                cost = (size, a, b)
                if best is None or cost < best[0]:
                    best = (cost, a, b, keep)
            if best is None:  # disconnected: outer product of the two smallest
                a, b = 0, 1
                keep = list(work[0].inds + work[1].inds)
            else:
                _, a, b, keep = best
            t = _letters(work[a].inds + work[b].inds)
            res = np.einsum(work[a].data, [t[i] for i in work[a].inds], work[b].data, [t[i] for i in work[b].inds],
                            [t[i] for i in keep], optimize=True)
            new = Tensor(res, keep)
            for x in (work[a], work[b]):
                for i in set(x.inds):
                    holders[i] -= 1
            for i in set(keep):
                holders[i] = holders.get(i, 0) + 1
            work = [x for k, x in enumerate(work) if k not in (a, b)] + [new]
        last = work[0]
        summed = [i for i in last.inds if i not in opened]
        if summed:
            last = contract_many([last], summed)
        return last

    def __len__(self):
        return len(self.tensors)
