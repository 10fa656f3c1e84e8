"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the MPS hot
path of Qrochet.jl, following `/root/reference/src/Ansatz/Chain.jl`,
`src/Ansatz.jl`, `src/Quantum.jl`, `src/Ansatz/Dense.jl` function by function.

PARITY STATUS:
  * The reference holds NO golden vectors / fixtures / seeds.  Its property tests
    (`test/Ansatz/Chain_test.jl:189-394`) pin truncate!, rand, canonize_site!,
    canonize!, mixed_canonize!, normalize!, adjoint; those are restated in
    `tests/test_oracle_properties.py` and this oracle passes them.
  * evolve!/expect/overlap have no reference test (`Chain_test.jl:396` is a TODO):
    PARITY UNPINNED by the reference for those; this oracle is instead checked
    against a dense state-vector simulation (`oracle/statevector.py`).
  * The reference (Julia) cannot run in this image; numerics use the same LAPACK
    drivers (zgesdd / zgeqrf / zgemm via SciPy/OpenBLAS).

Only `tests/`, `__graft_entry__.smoke()` and bench.py's cpu_baseline /
`--impl reference` legs may import this module.
"""
from __future__ import annotations

import numpy as np

from .tenet import Tensor, TensorNetwork, contract, gensym, nextindex


class MissingSchmidtCoefficientsException(Exception):
    """src/Ansatz.jl:91-99."""


def site(i: int, dual: bool = False):
    """`Site(i; dual)` (src/Quantum.jl:10-15)."""
    return (int(i), bool(dual))


class Quantum:
    """`Quantum` = TensorNetwork + Dict{Site,Symbol} (src/Quantum.jl:54-76)."""

    def __init__(self, tn: TensorNetwork, sites: dict):
        open_inds = set(tn.inds("open"))
        all_inds = set(tn.inds())
        for index in dict(sites).values():                      # src/Quantum.jl:61-67
            if index not in all_inds:
                raise RuntimeError(f"Index {index} not found in TensorNetwork")
            if index not in open_inds:
                raise RuntimeError(f"Index {index} must be open")
        self.tn = tn
        self.sites = dict(sites)

    def socket_type(self):
        """`socket(q)` (src/Quantum.jl): Scalar, State (dual when it only has inputs) or Operator."""
        ni, no = len(self.inputs()), len(self.outputs())
        if ni == 0 and no == 0:
            return "scalar"
        if ni == 0:
            return "state"
        if no == 0:
            return "state'"
        return "operator"

    def copy(self):
        return Quantum(self.tn.copy(), self.sites)

    def deepcopy(self):
        return Quantum(self.tn.deepcopy(), self.sites)

    def inputs(self):
        return sorted(s for s in self.sites if s[1])

    def outputs(self):
        return sorted(s for s in self.sites if not s[1])

    def nlanes(self):
        return len({s[0] for s in self.sites})

    def ind_at(self, s):
        """`inds(q; at=site)` (src/Quantum.jl:253-261)."""
        return self.sites[s]

    def tensor_at(self, s) -> Tensor:
        """`tensors(q; at=site)` (src/Quantum.jl:270-278): the tensor holding the site's index.
        (With a Λ-free physical index there is exactly one.)"""
        hits = self.tn.intersecting(self.sites[s])
        assert len(hits) == 1
        return hits[0]

    def adjoint(self):
        """src/Quantum.jl:100-115: conj + swap input/output + prime all virtual indices."""
        sites = {(s[0], not s[1]): i for s, i in self.sites.items()}
        tn = self.tn.conj()
        phys = set(sites.values())
        tn.replace_inds({i: i + "'" for i in tn.inds() if i not in phys})
        return Quantum(tn, sites)


class Product(Quantum):
    """`Product` ansatz (src/Ansatz/Product.jl): one tensor per lane, no inner indices.  Vectors give a State
    (:24-33), matrices an Operator with array dims (output, input) = inds [symbols[i+n], symbols[i]] (:35-46)."""

    def __init__(self, arrays=None, _q: Quantum | None = None):
        if _q is not None:
            assert not _q.tn.inds("inner"), "Product ansatz must not have inner indices"   # Product.jl:15
            super().__init__(_q.tn, _q.sites)
            return
        arrays = [np.asarray(a) for a in arrays]
        n = len(arrays)
        if all(a.ndim == 1 for a in arrays):
            sym = [nextindex() for _ in range(n)]
            tensors = [Tensor(a, [sym[k]]) for k, a in enumerate(arrays)]
            sites = {site(k + 1): sym[k] for k in range(n)}
        elif all(a.ndim == 2 for a in arrays):
            sym = [nextindex() for _ in range(2 * n)]
            tensors = [Tensor(a, [sym[k + n], sym[k]]) for k, a in enumerate(arrays)]
            sites = {site(k + 1, True): sym[k] for k in range(n)}
            sites.update({site(k + 1): sym[k + n] for k in range(n)})
        else:
            raise TypeError("Product takes a list of vectors (State) or of matrices (Operator)")
        super().__init__(TensorNetwork(tensors), sites)

    @staticmethod
    def zeros(n, p=2, dtype=bool):
        """`zeros(Product, n; p, eltype)` (Product.jl:48-50): |0...0>."""
        v = np.zeros(p, dtype=dtype)
        v[0] = 1
        return Product([v.copy() for _ in range(n)])

    @staticmethod
    def ones(n, p=2, dtype=bool):
        """`ones(Product, n; p, eltype)` (Product.jl:52-58): |1...1>."""
        v = np.zeros(p, dtype=dtype)
        v[1] = 1
        return Product([v.copy() for _ in range(n)])

    def copy(self):
        return Product(_q=Quantum(self.tn.deepcopy(), self.sites))

    def norm(self, p=2):
        """Product.jl:60-65 as written: (prod_i ||t_i||_p)^(1/p) -- for p = 2 the SQUARE ROOT of the product of the
        site norms (a defect of the reference: the norm of a product state is the plain product; replicated
        knowingly, it is exact for normalised sites, which is all the reference tests)."""
        prod = 1.0
        for t in self.tn.tensors:
            prod *= np.linalg.norm(np.ravel(t.data), p)
        return prod ** (1.0 / p)

    def opnorm(self, p=2):
        """Product.jl:67-72 (operators only), same outer exponent."""
        assert self.socket_type() == "operator"
        prod = 1.0
        for t in self.tn.tensors:
            prod *= np.linalg.norm(t.data, p)      # matrix p-norm: spectral norm for p = 2
        return prod ** (1.0 / p)

    def normalize_(self, p=2):
        """`normalize!` (Product.jl:74-80): every site tensor to unit p-norm."""
        for t in self.tn.tensors:
            t.data = t.data / np.linalg.norm(np.ravel(t.data), p)
        return self

    def overlap(self, other: "Product"):
        """`overlap(a::Product, b::Product)` (Product.jl:82-90): prod_i dot(a_i, conj(b_i)) -- Julia's `dot`
        conjugates its first argument, so this is prod_i sum_s conj(a_i[s]) conj(b_i[s])."""
        assert self.socket_type() == "state" and other.socket_type() == "state"
        assert set(self.sites) == set(other.sites), "Ansatzes must have the same sites"
        out = 1.0 + 0.0j
        for s in sorted(self.sites):
            out *= np.vdot(self.tensor_at(s).data, np.conj(other.tensor_at(s).data))
        return out


class Dense(Quantum):
    """Gate container, `Dense(Operator(), array; sites)` (src/Ansatz/Dense.jl:21-34)."""

    def __init__(self, array, sites):
        array = np.asarray(array)
        assert array.ndim == len(sites) and all(d > 1 for d in array.shape)
        inds = [nextindex() for _ in sites]
        super().__init__(TensorNetwork([Tensor(array, inds)]), dict(zip(sites, inds)))

    def copy(self):
        q = Quantum(self.tn.copy(), self.sites)
        q.__class__ = Dense
        return q


def gate(matrix, sites):
    """Helper: a k-site operator from a p^k x p^k matrix acting as |out><in|, array dims
    (o_1..o_k, i_1..i_k), first lane = fastest bit."""
    k = len(sites)
    matrix = np.asarray(matrix)
    p = round(matrix.shape[0] ** (1.0 / k))
    arr = np.reshape(matrix, (p,) * (2 * k), order="F")
    return Dense(arr, [site(s) for s in sites] + [site(s, True) for s in sites])


def _chain_labels(shapes, order, boundary, socket):
    """Index labels of the site tensors of a Chain, one list per site, plus the site map -- the four constructors
    `Chain(::State|::Operator, ::Open|::Periodic, arrays; order)` (Chain.jl:36-62, 64-100, 102-131, 133-172):
    site i holds o_i (and i_i for an operator), its right bond r_i = b_i and its left bond l_i = b_{i-1}, the bond
    ring closed (b_n between site n and site 1) for Periodic and cut there for Open (first array without `l`, last
    without `r`)."""
    default = ("o", "l", "r") if socket == "state" else ("o", "i", "l", "r")
    order = tuple(default if order is None else order)
    if sorted(order) != sorted(default):
        raise ValueError(f"order must be a permutation of {default}")                 # Chain.jl:39-40,68-69,105-106,137-138
    if boundary not in ("open", "periodic") or socket not in ("state", "operator"):
        raise ValueError("boundary must be 'open' or 'periodic', socket 'state' or 'operator'")
    n = len(shapes)
    full = len(default)
    out_ = [nextindex() for _ in range(n)]
    in_ = [nextindex() for _ in range(n)] if socket == "operator" else None
    bonds = [nextindex() for _ in range(n if boundary == "periodic" else n - 1)]
    labels = []
    for k in range(n):
        lab = {"o": out_[k], "i": in_[k] if in_ else None, "r": bonds[k % n] if (boundary == "periodic" or k < n - 1) else None,
               "l": bonds[(k - 1) % n] if (boundary == "periodic" or k > 0) else None}
        inds = [lab[c] for c in order if lab[c] is not None]
        want = full if boundary == "periodic" else full - (k == 0) - (k == n - 1)
        assert len(shapes[k]) == want == len(inds), f"array {k + 1} must have {want} dimensions"  # Chain.jl:37,65-67,103,134-136
        labels.append(inds)
    sites = {site(k + 1): out_[k] for k in range(n)}
    if in_:
        sites.update({site(k + 1, True): in_[k] for k in range(n)})
    return labels, sites


class Chain(Quantum):
    """`Chain(socket, boundary, arrays; order)`: MPS / MPO with open or periodic boundary (Chain.jl:6-31, 36-172).
    The default (`Chain(arrays)`) is the open-boundary MPS the algorithms of Chain.jl work on."""

    def __init__(self, arrays=None, order=None, _q: Quantum | None = None, boundary="open", socket="state"):
        self.boundary, self.socket = boundary, socket
        if _q is not None:
            super().__init__(_q.tn, _q.sites)
            return
        arrays = [np.asarray(a) for a in arrays]
        labels, sites = _chain_labels([a.shape for a in arrays], order, boundary, socket)
        super().__init__(TensorNetwork([Tensor(a, inds) for a, inds in zip(arrays, labels)]), sites)

    # -- bookkeeping ---------------------------------------------------------------------
    def copy(self):
        return Chain(_q=Quantum(self.tn.copy(), self.sites), boundary=self.boundary, socket=self.socket)

    def deepcopy(self):
        return Chain(_q=Quantum(self.tn.deepcopy(), self.sites), boundary=self.boundary, socket=self.socket)

    def adjoint(self):
        """`Chain(adjoint(Quantum(chain)), boundary(chain))` (Chain.jl:204)."""
        return Chain(_q=super().adjoint(), boundary=self.boundary, socket=self.socket)

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

    def nsites(self):
        """`nsites` = number of sites in the site map (2n for an operator; Quantum.jl)."""
        return len(self.sites)

    def bond_ind(self, s1, s2):
        """`inds(tn; bond=(s1,s2))` (Ansatz.jl:58-68)."""
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
        """`tensors(tn; between=(s1,s2))` (Ansatz.jl:78-89): `tn[bond]`, i.e. the tensor whose
        index set is exactly {bond}; `None` if the two sites do not share an index."""
        b = self.bond_ind(s1, s2)
        if b is None:
            return None
        try:
            return self.tn.select([b])
        except KeyError:
            return None

    # -- Chain.jl:309-333 ----------------------------------------------------------------
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

    # -- Chain.jl:335-376 ----------------------------------------------------------------
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

    # -- Chain.jl:390-422 ----------------------------------------------------------------
    def truncate(self, bond, threshold=None, maxdim=None):
        vind = self.rightindex(bond[0])
        if vind != self.leftindex(bond[1]):
            raise ValueError(f"Invalid bond {bond}")
        if vind not in self.tn.inds("hyper"):
            raise MissingSchmidtCoefficientsException(bond)
        spectrum = self.tn.select([vind]).data
        size = self.tn.size(vind)
        extent = range(min(size, maxdim)) if maxdim is not None else range(size)
        if threshold is None:
            threshold = 1e-16
        keep = [i for i in extent if abs(spectrum[i]) > threshold]
        self.tn.slice_(vind, keep)
        return self

    # -- Chain.jl:424-458 ----------------------------------------------------------------
    def _gram(self, s, keep_ind):
        t = self.tensor_at(s)
        if keep_ind is None:
            keep_ind = gensym("dummy")
            t = Tensor(t.data[..., None], t.inds + (keep_ind,))
        g = contract(t, t.conj().replace({keep_ind: gensym("new")}))
        return g.data

    def isleftcanonical(self, s, atol=1e-12):
        g = self._gram(s, self.rightindex(s))
        return bool(np.allclose(g, np.eye(g.shape[0]), atol=atol, rtol=0))

    def isrightcanonical(self, s, atol=1e-12):
        g = self._gram(s, self.leftindex(s))
        return bool(np.allclose(g, np.eye(g.shape[0]), atol=atol, rtol=0))

    # -- Chain.jl:469-497 ----------------------------------------------------------------
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
            inv = _pinv_diag(lam.data, 1e-64)
            self.tn.replace_tensor(a, contract(a, Tensor(inv, lam.inds), dims=()))
            self.tn.push(lam)
        return self

    # -- Chain.jl:509-536 ----------------------------------------------------------------
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
        self.tn.replace_tensor(lam, Tensor(lam.data / np.linalg.norm(lam.data), lam.inds))
        return self

    # -- Chain.jl:543-603 ----------------------------------------------------------------
    def evolve(self, g: Dense, threshold=None, maxdim=None, iscanonical=False, renormalize=False):
        ins, outs = g.inputs(), g.outputs()
        if not ins or not outs:
            raise ValueError("Gate must be an operator")
        if {(s[0], False) for s in ins} != set(outs):
            raise ValueError("Gate inputs and outputs must be the same")
        if not {(s[0], False) for s in ins} <= set(self.outputs()):
            raise ValueError("Gate inputs must be a subset of the TN sites")
        nl = g.nlanes()
        if nl == 1:
            self._evolve_1site(g)
        elif nl == 2:
            ids = sorted(s[0] for s in ins)
            if ids != list(range(ids[0], ids[-1] + 1)):
                raise ValueError("Gate lanes must be contiguous")
            self._evolve_2site(g, threshold, maxdim, iscanonical, renormalize)
        else:
            raise ValueError(f"Invalid number of lanes {nl}, maximum is 2")
        return self

    def _evolve_1site(self, g: Dense):
        g = g.copy()
        tmp = gensym("tmp")
        (gin,) = g.inputs()
        target = (gin[0], False)
        phys = self.sites[target]
        self.tn.replace_inds({phys: tmp})
        g.tn.replace_inds({g.sites[gin]: tmp})
        (gout,) = g.outputs()
        g.tn.replace_inds({g.sites[gout]: phys})
        self.tn.merge_(g.tn)
        self.tn.contract_index(tmp)

    # -- Chain.jl:606-661 ----------------------------------------------------------------
    def _evolve_2site(self, g: Dense, threshold, maxdim, iscanonical, renormalize):
        g = g.copy()
        g.sites = dict(g.sites)
        sitel, siter = sorted(g.outputs())
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
        for s in g.inputs():
            tmp = gensym("tmp")
            ren_q[self.sites[(s[0], False)]] = tmp
            ren_g[g.sites[s]] = tmp
        # qtn physical indices and gate inputs become the contracting indices ...
        self.tn.replace_inds(ren_q)
        g.tn.replace_inds(ren_g)
        # ... and the gate outputs take over the names in the site map
        g.tn.replace_inds({g.sites[s]: self.sites[s] for s in g.outputs()})
        self.tn.merge_(g.tn)
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
                self.tn.replace_tensor(lam, Tensor(lam.data / np.linalg.norm(lam.data), lam.inds))
            elif renormalize:
                self.normalize(bond[0])
        return self

    # -- Chain.jl:669-685 ----------------------------------------------------------------
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

    # -- Chain.jl:693-722 ----------------------------------------------------------------
    def unpack_2sitewf(self, bond, left_inds, right_inds, virtualind):
        from .tenet import svd

        sitel, siter = bond
        n = self.nsites()
        # NB: after contract_2sitewf both sites map to θ, so `between` must look at the index
        lam_l = None if sitel[0] == 1 else self._lambda_on(left_inds[0])
        lam_r = None if siter[0] == n else self._lambda_on(right_inds[0])
        theta = self.tensor_at(sitel)
        U, s, Vt = svd(theta, left_inds, right_inds, virtualind)
        gl = U if lam_l is None else contract(U, Tensor(_pinv_diag(lam_l.data, 1e-32), lam_l.inds), dims=())
        gr = Vt if lam_r is None else contract(Tensor(_pinv_diag(lam_r.data, 1e-32), lam_r.inds), Vt, dims=())
        self.tn.delete(theta)
        self.tn.push(gl)
        self.tn.push(s)
        self.tn.push(gr)
        return self

    def _lambda_on(self, ind):
        try:
            return self.tn.select([ind])
        except KeyError:
            return None

    # -- Chain.jl:724-752, Ansatz.jl:101-109 -----------------------------------------------
    def expect(self, observables):
        phi = self.copy()
        for o in observables:
            phi.evolve(o)
        tn = phi.tn.merge(self.adjoint().tn)
        return tn.contract().data[()]

    def overlap(self, other: "Chain"):
        """<other|self>: `other` is the conjugated one (Chain.jl:740-748)."""
        b = other.copy()
        b.tn.replace_inds({b.sites[s]: self.sites[s] for s in self.outputs()})
        b.sites = {s: self.sites[s] for s in self.outputs()}
        tn = self.tn.merge(b.adjoint().tn)
        return tn.contract().data[()]

    def norm(self):
        v = self.tn.merge(self.adjoint().tn).contract().data[()]
        return abs(np.sqrt(v))

    # -- helpers for the tests ---------------------------------------------------------------
    def lambdas(self):
        """Λ vector on each bond (None where absent), bond k between sites k and k+1 (1-based)."""
        out = []
        for k in range(1, self.nsites()):
            lam = self.lambda_between(site(k), site(k + 1))
            out.append(None if lam is None else lam.data)
        return out

    def to_dense(self):
        """Full state vector with qubit 1 the fastest index (column-major over sites)."""
        t = self.tn.contract()
        t = t.permute([self.sites[site(k + 1)] for k in range(self.nsites())])
        return np.reshape(t.data, -1, order="F")


def _pinv_diag(lam, atol):
    """`diag(pinv(Diagonal(λ), atol=atol))` (Chain.jl:491,710,713)."""
    lam = np.asarray(lam)
    out = np.zeros_like(lam)
    mask = np.abs(lam) > atol
    out[mask] = 1.0 / lam[mask]
    return out


def gramschmidt_rows(a):
    """`Muscle.gramschmidt!` [ext]: orthonormalise the ROWS of `a` (classical GS, in order)."""
    a = np.array(a)
    for i in range(a.shape[0]):
        for _ in range(2):
            if i:
                a[i] -= (a[:i].conj() @ a[i]) @ a[:i]
        a[i] /= np.linalg.norm(a[i])
    return a


def rand_mps_arrays(rng, n, chi, p=2, dtype=np.complex128, fast=False):
    """Arrays of `rand(Chain, Open, State; n, χ, p, eltype)` (Chain.jl:223-256), order (o,l,r).
    Entries U[0,1) (+ i U[0,1)) from `rng` (Julia's Xoshiro stream cannot be reproduced).
    `fast=True` replaces the O(χ³) Gram-Schmidt loop by a QR with the same row-space property
    (for the big bench shapes only; untimed set-up)."""
    arrays = []
    for i in range(1, n + 1):
        after_mid = i > n // 2
        j = (n + 1 - abs(2 * i - n - 1)) // 2
        chil, chir = min(chi, p ** (j - 1)), min(chi, p ** j)
        if n % 2 == 1 and i == n // 2 + 1:
            chil, chir = chil, chil
        elif after_mid:
            chil, chir = chir, chil
        if i == 1:
            chil, chir = chir, 1
        a = rng.random((chil, chir * p))
        if np.issubdtype(dtype, np.complexfloating):
            a = a + 1j * rng.random((chil, chir * p))
        a = a.astype(dtype)
        if fast:
            q, _ = np.linalg.qr(a.conj().T)
            a = q.conj().T.copy()
        else:
            a = gramschmidt_rows(a)
        a = np.reshape(a, (chil, chir, p), order="F")
        arrays.append(np.transpose(a, (2, 0, 1)))
    arrays[0] = np.reshape(arrays[0], (p, p), order="F")
    arrays[-1] = np.reshape(arrays[-1], (p, p), order="F")
    arrays[0] = arrays[0] / np.sqrt(p)
    return arrays


def rand_mps(rng, n, chi, p=2, dtype=np.complex128, fast=False) -> Chain:
    return Chain(rand_mps_arrays(rng, n, chi, p, dtype, fast))


def rand_mpo_arrays(rng, n, chi, p=2, dtype=np.float64):
    """Arrays of `rand(Chain, Open, Operator; n, χ, p, eltype)` (Chain.jl:260-297), default order (o,i,l,r):
    per site a random χl x (χr p²) matrix (first site χr x p², last χl x p²) with its ROWS orthonormalised
    (`Muscle.gramschmidt!`), reshaped and permuted to (p, p, χl, χr) / (p, p, χ); bond b has
    min(χ, p^(2b), p^(2(n-b))); site 1 divided by sqrt(min(χ, p²)) => Frobenius norm 1.  eltype defaults to Float64
    as in the reference (:264)."""
    arrays = []
    for i in range(1, n + 1):
        after_mid = i > n // 2
        j = (n + 1 - abs(2 * i - n - 1)) // 2
        chil, chir = min(chi, p ** (2 * (j - 1))), min(chi, p ** (2 * j))
        if n % 2 == 1 and i == n // 2 + 1:
            chil, chir = chil, chil
        elif after_mid:
            chil, chir = chir, chil
        shape = (chir, p, p) if i == 1 else ((chil, p, p) if i == n else (chil, chir, p, p))
        cols = int(np.prod(shape[1:]))
        a = rng.random((shape[0], cols))
        if np.issubdtype(dtype, np.complexfloating):
            a = a + 1j * rng.random((shape[0], cols))
        a = np.reshape(gramschmidt_rows(a.astype(dtype)), shape, order="F")
        arrays.append(np.transpose(a, (1, 2, 0)) if i in (1, n) else np.transpose(a, (2, 3, 0, 1)))
    arrays[0] = arrays[0] / np.sqrt(min(chi, p * p))
    return arrays


def rand_mpo(rng, n, chi, p=2, dtype=np.float64) -> Chain:
    return Chain(rand_mpo_arrays(rng, n, chi, p, dtype), socket="operator")


def haar_unitary(rng, d=4):
    """Haar-random d x d unitary: QR of complex Ginibre, phases fixed."""
    z = (rng.standard_normal((d, d)) + 1j * rng.standard_normal((d, d))) / np.sqrt(2)
    q, r = np.linalg.qr(z)
    ph = np.diag(r) / np.abs(np.diag(r))
    return q * ph


# ---- MPO x MPS: no function exists in the reference ------------------------------------
# Only `MPO(arrays)` (Chain.jl:133-172, default order (o, i, l, r), :34), `rand` MPO (:260-297) and the generic
# `merge(::Quantum, ::Quantum)` + `contract` exist; "MPO application with truncation" and "expect with an MPO" must
# be COMPOSED from them.  PARITY UNPINNED by the reference: the composition below is our definition, checked
# against dense linear algebra (tests/test_oracle_properties.py).
def heisenberg_mpo_arrays(n, J=1.0, h=0.0):
    """Spin-1/2 Heisenberg chain H = J sum S_k.S_{k+1} + h sum Sz_k as an MPO with D = 5, arrays (o, i, l, r);
    first site (o, i, r), last site (o, i, l)."""
    sz = np.diag([0.5, -0.5]).astype(np.complex128)
    sp = np.array([[0, 1], [0, 0]], dtype=np.complex128)
    sm = sp.T.copy()
    one = np.eye(2, dtype=np.complex128)
    w = np.zeros((5, 5, 2, 2), dtype=np.complex128)  # (l, r, o, i)
    w[0, 0] = one
    w[1, 0] = sp
    w[2, 0] = sm
    w[3, 0] = sz
    w[4, 0] = h * sz
    w[4, 1] = 0.5 * J * sm
    w[4, 2] = 0.5 * J * sp
    w[4, 3] = J * sz
    w[4, 4] = one
    bulk = np.transpose(w, (2, 3, 0, 1))  # (o, i, l, r)
    arrays = []
    for k in range(n):
        if k == 0:
            arrays.append(bulk[:, :, 4, :].copy())      # (o, i, r): last row
        elif k == n - 1:
            arrays.append(bulk[:, :, :, 0].copy())      # (o, i, l): first column
        else:
            arrays.append(bulk.copy())
    return arrays


def mpo_to_dense(arrays):
    """Dense operator of an open MPO given as (o, i, l, r) arrays; site 1 = fastest index."""
    n = len(arrays)
    t = arrays[0][:, :, None, :]  # (o, i, l=1, r)
    cur = np.transpose(t, (2, 0, 1, 3))[0]  # (o, i, r)
    cur = cur.reshape(2, 2, -1)
    for k in range(1, n):
        a = arrays[k] if k < n - 1 else arrays[k][:, :, :, None]
        # cur (O, I, r) x a (o, i, l=r, r') -> (O o, I i, r') with the new site slower
        cur = np.einsum("OIr,oirs->OoIis", cur, a)
        s = cur.shape
        cur = cur.reshape(s[0], s[1], s[2], s[3], s[4])
        cur = np.reshape(np.transpose(cur, (0, 1, 2, 3, 4)), (s[0], s[1], s[2], s[3], s[4]))
        cur = np.reshape(cur, (s[0] * s[1], s[2] * s[3], s[4]), order="C")
        # row index (O, o) in C order has o fastest -> we want the OLD sites fastest: swap
        cur = np.reshape(np.transpose(np.reshape(cur, (s[0], s[1], s[2], s[3], s[4])), (1, 0, 3, 2, 4)),
                         (s[1] * s[0], s[3] * s[2], s[4]), order="C")
    return cur[:, :, 0]


def apply_mpo_arrays(mps_arrays, mpo_arrays):
    """Site-wise `contract(merge(Quantum(ψ), Quantum(H)))` over the physical index: new MPS arrays (o, l, r) with
    fused bonds (ψ bond fastest, MPO bond slower)."""
    n = len(mps_arrays)
    out = []
    for k in range(n):
        a = np.asarray(mps_arrays[k])
        w = np.asarray(mpo_arrays[k])
        if k == 0:
            a = a[:, None, :]
            w = w[:, :, None, :]
        if k == n - 1:
            a = a[:, :, None] if a.ndim == 2 else a
            w = w[:, :, :, None] if w.ndim == 3 else w
        if k == 0 and a.ndim == 2:
            a = a[:, None, :]
        t = np.einsum("oiwx,ilr->olwrx", w, a)  # (o, l, w, r, x)
        o, l, wl, r, wr = t.shape
        t = np.reshape(t, (o, l * wl, r * wr), order="F")
        if k == 0:
            t = t[:, 0, :]
        if k == n - 1:
            t = t[:, :, 0] if t.ndim == 3 else t
        out.append(t)
    return out


def compress(chain: "Chain", maxdim=None, threshold=None):
    """`canonize!` (Chain.jl:469-497) with `truncate!` (:390-422) applied to each bond right after its SVD: the
    composition the reference's user writes to truncate an MPO-applied state."""
    n = chain.nsites()
    lams = []
    for i in range(n, 1, -1):
        chain.canonize_site(site(i), "left", "qr")
    for i in range(1, n):
        chain.canonize_site(site(i), "right", "svd")
        if maxdim is not None or threshold is not None:
            chain.truncate((site(i), site(i + 1)), threshold=threshold, maxdim=maxdim)
        lam = chain.lambda_between(site(i), site(i + 1))
        chain.tn.pop(lam)
        a = chain.tensor_at(site(i + 1))
        chain.tn.replace_tensor(a, contract(a, lam, dims=()))
        lams.append(lam)
    for i in range(2, n + 1):
        lam = lams[i - 2]
        a = chain.tensor_at(site(i))
        chain.tn.replace_tensor(a, contract(a, Tensor(_pinv_diag(lam.data, 1e-64), lam.inds), dims=()))
        chain.tn.push(lam)
    return chain


def expect_mpo(chain: "Chain", mpo_arrays):
    """<ψ|H|ψ> = contract(merge(ψ, H, ψ')) (un-normalised), composed as overlap(Hψ, ψ)."""
    dense_arrays = []
    # bring ψ to plain (o, l, r) arrays by absorbing any Λ to the right
    c = chain.copy()
    n = c.nsites()
    for k in range(1, n):
        if c.lambda_between(site(k), site(k + 1)) is not None:
            c.contract_between(site(k), site(k + 1), direction="left")
    for k in range(1, n + 1):
        t = c.tensor_at(site(k))
        order = [c.sites[site(k)]]
        if k > 1:
            order.append(c.leftindex(site(k)))
        if k < n:
            order.append(c.rightindex(site(k)))
        dense_arrays.append(t.permute(order).data)
    hpsi = Chain(apply_mpo_arrays(dense_arrays, mpo_arrays))
    return hpsi.overlap(Chain(dense_arrays))


def chain_from_vidal(sites_lor, lambdas) -> "Chain":
    """Test helper: an open-boundary MPS from site arrays in the device's private (l, o, r) layout (edge bonds of
    size 1 kept) plus the Schmidt vectors sitting on its bonds (`None` where absent) -- the labelled network
    `canonize!` / Vidal `evolve!` leave behind (Λ on a hyperindex, Chain.jl:488-494).  Lets a parity test start the
    oracle from exactly the state the device holds."""
    n = len(sites_lor)
    arrays = []
    for k, a in enumerate(sites_lor):
        a = np.transpose(np.asarray(a), (1, 0, 2))  # -> (o, l, r), the reference's default order (Chain.jl:33)
        if k == 0:
            a = a[:, 0, :]
        if k == n - 1:
            a = a[..., 0]
        arrays.append(np.array(a))
    q = Chain(arrays)
    for b, lam in enumerate(lambdas):
        if lam is not None:
            q.tn.push(Tensor(np.array(lam, dtype=np.float64), [q.bond_ind(site(b + 1), site(b + 2))]))
    return q
