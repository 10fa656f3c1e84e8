"""Label-driven mirror of the reference's own code path on DEVICE tensors.

In Julia, after `adapt(B200Array, ψ)` the unmodified `src/Ansatz/Chain.jl` keeps running: it manipulates labelled
`Tenet.Tensor`s and every numeric call (`contract`, `svd!`, `qr!`, `slice!`, `conj`) dispatches on the device array
type and lands in libqrochet_b200.so (Tenet-level entry points of include/qrochet_b200.h).  This module is that
path in Python: `Tensor` = device array + index labels, `TensorNetwork` = bag of tensors with hyper-index
semantics, `Chain` = the MPS algorithms of Chain.jl transcribed line by line (each method cites its lines) -- no
NumPy arithmetic anywhere, only C-ABI calls.  The fused `B200MPS` (mps.py) is the fast path; this is the
drop-in path, and the tests run both against the oracle."""
from __future__ import annotations

import itertools

import numpy as np

from . import device as dev
from ._capi import MissingSchmidtCoefficientsException

_counter = itertools.count()


def nextindex() -> str:
    """`Qrochet.nextindex` (src/Utils.jl:3-7)."""
    return f"i{next(_counter)}"


def gensym(tag="tmp") -> str:
    return f"##{tag}#{next(_counter)}"


def site(i, dual=False):
    """`Site(i; dual)` (src/Quantum.jl:10-15)."""
    return (int(i), bool(dual))


class Tensor:
    """`Tenet.Tensor(data::B200Array, inds)`; real vectors (Λ) are F64 device arrays."""

    __slots__ = ("data", "inds")

    def __init__(self, data: dev.DeviceArray, inds):
        self.data = data
        self.inds = tuple(inds)
        assert data.ndim == len(self.inds), (data.shape, self.inds)

    @property
    def shape(self):
        return self.data.shape

    def size(self, ind):
        return self.data.shape[self.inds.index(ind)]

    def replace(self, mapping):
        return Tensor(self.data, tuple(mapping.get(i, i) for i in self.inds))

    def conj(self):
        if self.data.dtype in (2, 3):  # real Λ vector
            return Tensor(self.data, self.inds)
        return Tensor(dev.conj(self.data), self.inds)

    def to_host(self):
        return self.data.to_host()


class _Modes:
    """Symbol <-> int32 mode-label map (lives on the Julia side of the boundary)."""

    def __init__(self):
        self.table = {}

    def __call__(self, inds):
        return [self.table.setdefault(i, len(self.table)) for i in inds]


def _is_vector(t: Tensor):
    return t.data.dtype in (2, 3)  # F64 / F32: a real Schmidt vector


def contract(a: Tensor, b: Tensor, dims=None) -> Tensor:
    """`Tenet.contract(a, b; dims)` (call sites Chain.jl:322,325,435,453,484,491,710,713): sums `dims` (default all
    shared indices), `dims=()` keeps them; output order (inds(a) ∪ inds(b)) \\ dims.  A real vector operand (Λ or
    pinv(Λ)) with its index kept is the mode-scale kernel, everything else the permutation-fused GEMM."""
    shared = [i for i in a.inds if i in b.inds]
    dims = shared if dims is None else [i for i in dims if i in shared]
    out = [i for i in a.inds if i not in dims] + [i for i in b.inds if i not in a.inds]
    for vec, ten in ((a, b), (b, a)):
        if _is_vector(vec) and not _is_vector(ten) and not dims and vec.inds[0] in ten.inds:
            scaled = dev.scale_mode(ten.data, ten.inds.index(vec.inds[0]), vec.data)
            res = Tensor(scaled, ten.inds)
            return res if tuple(out) == ten.inds else permute(res, out)
    if _is_vector(a) or _is_vector(b):
        raise NotImplementedError("contraction summing over a real vector's index")
    m = _Modes()
    res = dev.contract(a.data, m(a.inds), b.data, m(b.inds), m(out))
    return Tensor(res, out)


def permute(t: Tensor, inds) -> Tensor:
    inds = tuple(inds)
    return Tensor(dev.permute(t.data, [t.inds.index(i) for i in inds]), inds)


def _split(t: Tensor, left_inds, right_inds):
    left_inds = list(left_inds)
    right_inds = list(right_inds) if right_inds else [i for i in t.inds if i not in left_inds]
    if not left_inds:
        left_inds = [i for i in t.inds if i not in right_inds]
    order = [t.inds.index(i) for i in left_inds + right_inds]
    return left_inds, right_inds, order


def svd(t: Tensor, left_inds, right_inds, virtualind):
    """`LinearAlgebra.svd(::Tensor; left_inds, right_inds, virtualind)` (Chain.jl:365,645,705)."""
    left_inds, right_inds, order = _split(t, left_inds, right_inds)
    u, s, vc, _, _ = dev.svd(t.data, order, len(left_inds))
    return Tensor(u, left_inds + [virtualind]), Tensor(s, [virtualind]), Tensor(vc, right_inds + [virtualind])


def qr(t: Tensor, left_inds, right_inds, virtualind):
    """`LinearAlgebra.qr(::Tensor; ...)` (Chain.jl:367)."""
    left_inds, right_inds, order = _split(t, left_inds, right_inds)
    q, r = dev.qr(t.data, order, len(left_inds))
    return Tensor(q, left_inds + [virtualind]), Tensor(r, [virtualind] + right_inds)


class TensorNetwork:
    """`Tenet.TensorNetwork` restricted to what Chain.jl uses."""

    def __init__(self, tensors=()):
        self.tensors = list(tensors)

    def copy(self):
        return TensorNetwork(self.tensors)

    def _count(self):
        c = {}
        for t in self.tensors:
            for i in t.inds:
                c[i] = c.get(i, 0) + 1
        return c

    def inds(self, set="all"):
        c = self._count()
        if set == "all":
            return list(c)
        lo = {"open": (1, 1), "inner": (2, 10 ** 9), "hyper": (3, 10 ** 9)}[set]
        return [i for i, n in c.items() if lo[0] <= n <= lo[1]]

    def size(self, ind):
        for t in self.tensors:
            if ind in t.inds:
                return t.size(ind)
        raise KeyError(ind)

    def intersecting(self, ind):
        return [t for t in self.tensors if ind in t.inds]

    def select(self, inds):
        want = frozenset(inds)
        hits = [t for t in self.tensors if frozenset(t.inds) == want]
        if len(hits) != 1:
            raise KeyError(sorted(want))
        return hits[0]

    def push(self, t):
        self.tensors.append(t)

    def delete(self, t):
        self.tensors = [x for x in self.tensors if x is not t]

    pop = delete

    def replace_tensor(self, old, new):
        self.tensors = [new if x is old else x for x in self.tensors]

    def replace_inds(self, mapping):
        self.tensors = [t.replace(mapping) if any(i in mapping for i in t.inds) else t for t in self.tensors]

    def merge(self, other):
        return TensorNetwork(self.tensors + other.tensors)

    def merge_(self, other):
        self.tensors += other.tensors
        return self

    def conj(self):
        return TensorNetwork([t.conj() for t in self.tensors])

    def contract_index(self, inds):
        """`contract!(tn, i)` (Chain.jl:372,602,616,636,682)."""
        inds = [inds] if isinstance(inds, str) else list(inds)
        hit = [t for t in self.tensors if any(i in t.inds for i in inds)]
        if not hit:
            return None
        rest = [t for t in self.tensors if not any(t is h for h in hit)]
        # order: scale vectors into their neighbours, then one GEMM summing the indices
        vecs = [t for t in hit if _is_vector(t)]
        dense = [t for t in hit if not _is_vector(t)]
        for v in vecs:
            k = next(k for k, d in enumerate(dense) if v.inds[0] in d.inds)
            dense[k] = contract(dense[k], v, dims=())
        res = dense[0]
        for d in dense[1:]:
            res = contract(res, d, dims=[i for i in inds if i in res.inds and i in d.inds])
        self.tensors = rest + [res]
        return res

    def slice_(self, ind, count):
        """`slice!(tn, ind, 1:count)` (Chain.jl:419)."""
        self.tensors = [Tensor(dev.slice_mode(t.data, t.inds.index(ind), count), t.inds) if ind in t.inds else t
                        for t in self.tensors]

    def svd_(self, left_inds, right_inds, virtualind):
        t = self.select(list(left_inds) + list(right_inds))
        U, S, Vt = svd(t, left_inds, right_inds, virtualind)
        self.delete(t)
        self.tensors += [U, S, Vt]

    def qr_(self, left_inds, right_inds, virtualind):
        t = self.select(list(left_inds) + list(right_inds))
        Q, R = qr(t, left_inds, right_inds, virtualind)
        self.delete(t)
        self.tensors += [Q, R]

    def contract(self) -> Tensor:
        """`contract(tn)`: Λ vectors are absorbed into a neighbour (scale), then a greedy pairwise contraction
        (smallest result first); an index is summed once no other tensor holds it."""
        opened = set(self.inds("open"))
        vecs = [t for t in self.tensors if _is_vector(t)]
        work = [t for t in self.tensors if not _is_vector(t)]
        for v in vecs:  # Λ on a hyper index: scale ONE holder, the index stays shared by the others
            k = next(k for k, d in enumerate(work) if v.inds[0] in d.inds)
            work[k] = contract(work[k], v, dims=())
        while len(work) > 1:
            best = None
            for a, b in itertools.combinations(range(len(work)), 2):
                ia, ib = work[a].inds, work[b].inds
                if not set(ia) & set(ib):
                    continue
                others = set()
                for k, t in enumerate(work):
                    if k not in (a, b):
                        others.update(t.inds)
                keep = [i for i in dict.fromkeys(ia + ib) if i in opened or i in others]
                size = 1
                for i in keep:
                    size *= work[a].size(i) if i in ia else work[b].size(i)
                if best is None or (size, a, b) < best[0]:
                    best = ((size, a, b), a, b, keep)
            if best is None:
                a, b = 0, 1
                keep = list(work[0].inds + work[1].inds)
            else:
                _, a, b, keep = best
            m = _Modes()
            res = dev.contract(work[a].data, m(work[a].inds), work[b].data, m(work[b].inds), m(keep))
            work = [x for k, x in enumerate(work) if k not in (a, b)] + [Tensor(res, keep)]
        return work[0]

    def __len__(self):
        return len(self.tensors)


def _site_labels(shapes, order, boundary, socket):
    """Index labels per site tensor + site map for the four reference constructors
    `Chain(::State|::Operator, ::Open|::Periodic, arrays; order)` (Chain.jl:36-62, 64-100, 102-131, 133-172)."""
    default = ("o", "l", "r") if socket == "state" else ("o", "i", "l", "r")
    order = tuple(default if order is None else order)
    if sorted(order) != sorted(default):
        raise ValueError(f"order must be a permutation of {default}")      # ArgumentError, Chain.jl:39-40,68-69,105-106
    if boundary not in ("open", "periodic") or socket not in ("state", "operator"):
        raise ValueError("boundary must be 'open' or 'periodic', socket 'state' or 'operator'")
    n = len(shapes)
    ring = boundary == "periodic"
    phys = [nextindex() for _ in range(n)]
    dual = [nextindex() for _ in range(n)] if socket == "operator" else None
    bonds = [nextindex() for _ in range(n if ring else n - 1)]
    labels = []
    for k in range(n):
        lab = {"o": phys[k], "i": dual[k] if dual else None,
               "l": bonds[(k - 1) % n] if (ring or k > 0) else None,
               "r": bonds[k % n] if (ring or k < n - 1) else None}
        inds = [lab[c] for c in order if lab[c] is not None]
        if len(shapes[k]) != len(inds):                                     # the @asserts of Chain.jl:37,65-67,103,134-136
            raise AssertionError(f"array {k + 1} must have {len(inds)} dimensions")
        labels.append(inds)
    sites = {site(k + 1): phys[k] for k in range(n)}
    if dual:
        sites.update({site(k + 1, True): dual[k] for k in range(n)})
    return labels, sites


class Chain:
    """`Chain(socket, boundary, arrays; order)` (Chain.jl:6-31, 36-172) with the arrays uploaded at construction
    (= `adapt(B200Array, ψ)`, ext/QrochetAdaptExt.jl:9).  The default is the open-boundary MPS the algorithms of
    Chain.jl work on; periodic chains and operators (MPO) carry the same bookkeeping and go through the generic
    network contraction (`norm`, `overlap`, `adjoint`)."""

    def __init__(self, ctx: dev.Context | None = None, arrays=None, order=None, _state=None, boundary="open",
                 socket="state"):
        self.boundary, self.socket = boundary, socket
        if _state is not None:
            self.ctx, self.tn, self.sites = _state
            return
        self.ctx = ctx
        host = []
        for a in arrays:
            a = np.asarray(a)
            if a.dtype != np.complex64:  # ComplexF32 chains stay ComplexF32 on the device (native float2 tensors)
                a = a.astype(np.complex128)
            host.append(a)
        labels, self.sites = _site_labels([a.shape for a in host], order, boundary, socket)
        self.tn = TensorNetwork([Tensor(ctx.array(a), inds) for a, inds in zip(host, labels)])

    @property
    def eltype(self):
        """Element type of the site tensors (ComplexF64, or ComplexF32 when built from complex64 arrays)."""
        return np.complex64 if any(t.data.dtype == 1 for t in self.tn.tensors) else np.complex128

    # ---- bookkeeping (src/Quantum.jl, src/Ansatz.jl) ----
    def copy(self):
        return Chain(_state=(self.ctx, self.tn.copy(), dict(self.sites)), boundary=self.boundary, socket=self.socket)

    def nsites(self):
        return len(self.sites)

    def nlanes(self):
        return len({s[0] for s in self.sites})

    def outputs(self):
        return sorted(s for s in self.sites if not s[1])

    def inputs(self):
        return sorted(s for s in self.sites if s[1])

    def leftsite(self, s):
        """Chain.jl:185-188."""
        if self.boundary == "periodic":
            return ((s[0] - 2) % self.nlanes() + 1, s[1])
        return (s[0] - 1, s[1]) if 2 <= s[0] <= self.nlanes() else None

    def rightsite(self, s):
        """Chain.jl:190-193."""
        if self.boundary == "periodic":
            return (s[0] % self.nlanes() + 1, s[1])
        return (s[0] + 1, s[1]) if 1 <= s[0] <= self.nlanes() - 1 else None

    def tensor_at(self, s):
        (t,) = self.tn.intersecting(self.sites[s])
        return t

    def bond_ind(self, s1, s2):
        a, b = self.tensor_at(s1), self.tensor_at(s2)
        common = [i for i in a.inds if i in b.inds]
        if not common:
            return None
        (only,) = common
        return only

    def leftindex(self, s):
        """Chain.jl:195-197."""
        ls = self.leftsite(s)
        return None if ls is None else self.bond_ind(s, ls)

    def rightindex(self, s):
        """Chain.jl:199-202."""
        rs = self.rightsite(s)
        return None if rs is None else self.bond_ind(s, rs)

    def lambda_between(self, s1, s2):
        b = self.bond_ind(s1, s2)
        if b is None:
            return None
        try:
            return self.tn.select([b])
        except KeyError:
            return None

    def adjoint(self):
        """src/Quantum.jl:100-115."""
        sites = {(s[0], not s[1]): i for s, i in self.sites.items()}
        tn = self.tn.conj()
        phys = set(sites.values())
        tn.replace_inds({i: i + "'" for i in tn.inds() if i not in phys})
        return Chain(_state=(self.ctx, tn, sites), boundary=self.boundary, socket=self.socket)  # Chain.jl:204

    def _pinv(self, lam: Tensor, atol):
        """`Tensor(diag(pinv(Diagonal(parent(Λ)), atol)), inds(Λ))` (Chain.jl:491,710,713): tiny host round trip of
        the Schmidt vector (the reference reads it element-wise on the host too, Chain.jl:402,416)."""
        v = lam.to_host()
        out = np.zeros_like(v)
        mask = np.abs(v) > atol
        out[mask] = 1.0 / v[mask]
        return Tensor(self.ctx.array(out), lam.inds)

    # ---- Chain.jl:309-333 ----
    def contract_between(self, s1, s2, direction="left", delete_lambda=True):
        lam = self.lambda_between(s1, s2)
        if lam is None:
            return self
        if direction == "right":
            g = self.tensor_at(s2)
            self.tn.replace_tensor(g, contract(g, lam, dims=()))
        elif direction == "left":
            g = self.tensor_at(s1)
            self.tn.replace_tensor(g, contract(lam, g, dims=()))
        else:
            raise ValueError(f"Unknown direction=:{direction}")
        if delete_lambda:
            self.tn.delete(lam)
        return self

    # ---- Chain.jl:335-376 ----
    def canonize_site(self, s, direction, method="qr"):
        n = self.nsites()
        left_inds, right_inds = [], []
        if direction == "left":
            if s == site(1):
                raise ValueError("Cannot right-canonize left-most tensor")
            right_inds.append(self.leftindex(s))
            if s != site(n):
                left_inds.append(self.rightindex(s))
            left_inds.append(self.sites[s])
        elif direction == "right":
            if s == site(n):
                raise ValueError("Cannot left-canonize right-most tensor")
            right_inds.append(self.rightindex(s))
            if s != site(1):
                left_inds.append(self.leftindex(s))
            left_inds.append(self.sites[s])
        else:
            raise ValueError(f"Unknown direction=:{direction}")
        (virtualind,) = right_inds
        tmp = gensym("tmp")
        if method == "svd":
            self.tn.svd_(left_inds, right_inds, tmp)
        elif method == "qr":
            self.tn.qr_(left_inds, right_inds, tmp)
        else:
            raise ValueError(f"Unknown factorization method=:{method}")
        self.tn.contract_index(virtualind)
        self.tn.replace_inds({tmp: virtualind})
        return self

    # ---- Chain.jl:390-422 ----
    def truncate(self, bond, threshold=None, maxdim=None):
        vind = self.rightindex(bond[0])
        if vind != self.leftindex(bond[1]):
            raise ValueError(f"Invalid bond {bond}")
        if vind not in self.tn.inds("hyper"):
            raise MissingSchmidtCoefficientsException(-4, f"Can't access the spectrum on bond {bond}")
        spectrum = self.tn.select([vind]).to_host()
        size = self.tn.size(vind)
        extent = range(min(size, maxdim)) if maxdim is not None else range(size)
        if threshold is None:
            threshold = 1e-16
        keep = [i for i in extent if abs(spectrum[i]) > threshold]
        assert keep == list(range(len(keep)))  # sorted spectrum: always a prefix
        self.tn.slice_(vind, len(keep))
        return self

    # ---- Chain.jl:424-458 ----
    def _gram(self, s, keep_ind):
        t = self.tensor_at(s)
        if keep_ind is None:
            keep_ind = gensym("dummy")
            t = Tensor(t.data.reshape(t.shape + (1,)), t.inds + (keep_ind,))
        g = contract(t, t.conj().replace({keep_ind: gensym("new")}))
        return g.to_host()

    def isleftcanonical(self, s, atol=1e-12):
        g = self._gram(s, self.rightindex(s))
        return bool(np.allclose(g, np.eye(g.shape[0]), atol=atol, rtol=0))

    def isrightcanonical(self, s, atol=1e-12):
        g = self._gram(s, self.leftindex(s))
        return bool(np.allclose(g, np.eye(g.shape[0]), atol=atol, rtol=0))

    # ---- Chain.jl:469-497 ----
    def canonize(self):
        n = self.nsites()
        lams = []
        for i in range(n, 1, -1):
            self.canonize_site(site(i), "left", "qr")
        for i in range(1, n):
            self.canonize_site(site(i), "right", "svd")
            lam = self.lambda_between(site(i), site(i + 1))
            self.tn.pop(lam)
            a = self.tensor_at(site(i + 1))
            self.tn.replace_tensor(a, contract(a, lam, dims=()))
            lams.append(lam)
        for i in range(2, n + 1):
            lam = lams[i - 2]
            a = self.tensor_at(site(i))
            self.tn.replace_tensor(a, contract(a, self._pinv(lam, 1e-64), dims=()))
            self.tn.push(lam)
        return self

    # ---- Chain.jl:509-536 ----
    def mixed_canonize(self, center):
        n = self.nsites()
        for i in range(1, center[0]):
            self.canonize_site(site(i), "right", "qr")
        for i in range(n, center[0], -1):
            self.canonize_site(site(i), "left", "qr")
        self.canonize_site(center, "left", "svd")
        return self

    def normalize(self, root):
        self.mixed_canonize(root)
        lam = self.lambda_between(site(root[0] - 1), root)
        nrm = dev.norm2(lam.data)
        self.tn.replace_tensor(lam, Tensor(dev.scale(lam.data.copy(), 1.0 / nrm), lam.inds))
        return self

    # ---- Chain.jl:543-661 (gate = array with dims (o_1.., i_1..), lanes 1-based) ----
    def evolve(self, gate_array, lanes, threshold=None, maxdim=None, iscanonical=False, renormalize=False):
        lanes = list(lanes)
        k = len(lanes)
        g_inds = [nextindex() for _ in range(2 * k)]
        g = Tensor(self.ctx.array(np.asarray(gate_array, dtype=self.eltype)), g_inds)
        g_sites = {**{site(l): g_inds[j] for j, l in enumerate(lanes)},
                   **{site(l, True): g_inds[k + j] for j, l in enumerate(lanes)}}
        if not {site(l) for l in lanes} <= set(self.outputs()):
            raise ValueError("Gate inputs must be a subset of the TN sites")
        if k == 1:
            return self._evolve_1site(g, g_sites, lanes[0])
        if k == 2:
            if sorted(lanes) != list(range(min(lanes), max(lanes) + 1)):
                raise ValueError("Gate lanes must be contiguous")
            return self._evolve_2site(g, g_sites, sorted(lanes), threshold, maxdim, iscanonical, renormalize)
        raise ValueError(f"Invalid number of lanes {k}, maximum is 2")

    def _evolve_1site(self, g, g_sites, lane):
        tmp = gensym("tmp")
        phys = self.sites[site(lane)]
        self.tn.replace_inds({phys: tmp})
        g = g.replace({g_sites[site(lane, True)]: tmp}).replace({g_sites[site(lane)]: phys})
        self.tn.push(g)
        self.tn.contract_index(tmp)
        return self

    def _evolve_2site(self, g, g_sites, lanes, threshold, maxdim, iscanonical, renormalize):
        sitel, siter = site(lanes[0]), site(lanes[1])
        bond = (sitel, siter)
        li, ri = self.leftindex(sitel), self.rightindex(siter)
        left_inds = [li] if li is not None else []
        right_inds = [ri] if ri is not None else []
        virtualind = self.bond_ind(sitel, siter)
        if iscanonical:
            self.contract_2sitewf(bond)
        else:
            self.tn.contract_index(virtualind)
        ren_q, ren_g = {}, {}
        for l in lanes:
            tmp = gensym("tmp")
            ren_q[self.sites[site(l)]] = tmp
            ren_g[g_sites[site(l, True)]] = tmp
        self.tn.replace_inds(ren_q)
        g = g.replace(ren_g).replace({g_sites[site(l)]: self.sites[site(l)] for l in lanes})
        self.tn.push(g)
        self.tn.contract_index(list(ren_q.values()))
        left_inds.append(self.sites[sitel])
        right_inds.append(self.sites[siter])
        if iscanonical:
            self.unpack_2sitewf(bond, left_inds, right_inds, virtualind)
        else:
            self.tn.svd_(left_inds, right_inds, virtualind)
        if threshold is not None or maxdim is not None:
            self.truncate(bond, threshold=threshold, maxdim=maxdim)
            if renormalize and iscanonical:
                lam = self.lambda_between(*bond)
                nrm = dev.norm2(lam.data)
                self.tn.replace_tensor(lam, Tensor(dev.scale(lam.data.copy(), 1.0 / nrm), lam.inds))
            elif renormalize:
                self.normalize(bond[0])
        return self

    # ---- Chain.jl:669-722 ----
    def contract_2sitewf(self, bond):
        sitel, siter = bond
        n = self.nsites()
        lam_l = None if sitel[0] == 1 else self.lambda_between(site(sitel[0] - 1), sitel)
        lam_r = None if sitel[0] == n - 1 else self.lambda_between(siter, site(siter[0] + 1))
        if lam_l is not None:
            self.contract_between(site(sitel[0] - 1), sitel, direction="right", delete_lambda=False)
        if lam_r is not None:
            self.contract_between(siter, site(siter[0] + 1), direction="left", delete_lambda=False)
        self.tn.contract_index(self.bond_ind(sitel, siter))
        return self

    def unpack_2sitewf(self, bond, left_inds, right_inds, virtualind):
        sitel, siter = bond
        n = self.nsites()

        def lam_on(ind):
            try:
                return self.tn.select([ind])
            except KeyError:
                return None
        lam_l = None if sitel[0] == 1 else lam_on(left_inds[0])
        lam_r = None if siter[0] == n else lam_on(right_inds[0])
        theta = self.tensor_at(sitel)
        U, s, Vt = svd(theta, left_inds, right_inds, virtualind)
        gl = U if lam_l is None else contract(U, self._pinv(lam_l, 1e-32), dims=())
        gr = Vt if lam_r is None else contract(self._pinv(lam_r, 1e-32), Vt, dims=())
        self.tn.delete(theta)
        self.tn.push(gl)
        self.tn.push(s)
        self.tn.push(gr)
        return self

    # ---- Chain.jl:724-752, Ansatz.jl:101-109 ----
    def expect(self, observables):
        """`observables`: list of (gate_array, lanes)."""
        phi = self.copy()
        for g, lanes in observables:
            phi.evolve(g, lanes)
        return complex(phi.tn.merge(self.adjoint().tn).contract().to_host()[()])

    def overlap(self, other: "Chain"):
        b = other.copy()
        b.tn.replace_inds({b.sites[s]: self.sites[s] for s in self.outputs()})
        b.sites = {s: self.sites[s] for s in self.outputs()}
        return complex(self.tn.merge(b.adjoint().tn).contract().to_host()[()])

    def norm(self):
        return abs(np.sqrt(complex(self.tn.merge(self.adjoint().tn).contract().to_host()[()])))

    # ---- helpers for tests ----
    def lambdas(self):
        out = []
        for k in range(1, self.nsites()):
            lam = self.lambda_between(site(k), site(k + 1))
            out.append(None if lam is None else lam.to_host())
        return out

    def to_dense(self):
        t = self.tn.contract()
        t = permute(t, [self.sites[site(k + 1)] for k in range(self.nsites())])
        return np.reshape(t.to_host(), -1, order="F")
