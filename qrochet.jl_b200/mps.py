"""`B200MPS`: the fused, device-resident open-boundary MPS (the "B200Chain wrapper" of SURVEY.md §8b) --
canonize!/mixed_canonize!/truncate!/evolve!/overlap/expect of /root/reference/src/Ansatz/Chain.jl run as
fused kernel chains inside libqrochet_b200.so.  Site arrays cross the boundary in the reference's
`defaultorder` (o, l, r) (Chain.jl:33) and are re-laid out to the private (l, o, r) once, at adapt time."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import check, lib
from .device import Context


_HOST_CODES = {np.dtype(np.complex128): capi.C128, np.dtype(np.complex64): capi.C64,
               np.dtype(np.float64): capi.F64, np.dtype(np.float32): capi.F32}


def _host_code(a: np.ndarray):
    """(array, element-type code) for the boundary: the four `eltype`s of `rand` (Chain.jl:226-227) cross as they are,
    anything else (integers, bool: `zeros(Product, n)`) is converted to ComplexF64 on the host."""
    if a.dtype not in _HOST_CODES:
        a = a.astype(np.complex128)
    return a, _HOST_CODES[a.dtype]


def _storage(arrays, dtype):
    """Storage type of the chain: ComplexF32 only when asked for, or when every site array is single precision."""
    if dtype is not None:
        dt = np.dtype(dtype)
        if dt not in (np.dtype(np.complex128), np.dtype(np.complex64)):
            raise ValueError("storage type must be complex128 or complex64 (real data is held as complex)")
        return capi.C64 if dt == np.dtype(np.complex64) else capi.C128
    single = all(np.asarray(a).dtype in (np.dtype(np.complex64), np.dtype(np.float32)) for a in arrays)
    return capi.C64 if single and len(arrays) else capi.C128


class B200MPS:
    def __init__(self, ctx: Context, arrays=None, order=("o", "l", "r"), dtype=None, _handle=None):
        """`arrays`: site arrays in the reference's `order` (default (o, l, r), Chain.jl:33) of any of the reference's
        element types (ComplexF64, ComplexF32, Float64, Float32; real chains are held as complex).  `dtype`: storage
        type in HBM, complex128 or complex64 (default: complex64 only if every array is single precision); the fused
        chains compute in FP64 either way."""
        self.ctx = ctx
        if _handle is not None:
            self.h = _handle
            return
        n = len(arrays)
        h = C.c_void_p()
        check(ctx.h, lib.qb200_mps_create_typed(ctx.h, n, _storage(arrays, dtype), C.byref(h)))
        self.h = h
        for k, a in enumerate(arrays):
            a, code = _host_code(np.asarray(a))
            labels = [c for c in order if not ((c == "l" and k == 0) or (c == "r" and k == n - 1))]
            assert a.ndim == len(labels), (k, a.shape, labels)
            # -> (l, o, r) with size-1 edge bonds
            axes = {c: i for i, c in enumerate(labels)}
            perm = [axes[c] for c in ("l", "o", "r") if c in axes]
            a = np.transpose(a, perm)
            if k == 0:
                a = a[None, ...]
            if k == n - 1:
                a = a[..., None]
            buf = np.asfortranarray(a)
            check(ctx.h, lib.qb200_mps_set_site_typed(ctx.h, self.h, k, code, buf.shape[0], buf.shape[1], buf.shape[2],
                                                      buf.ctypes.data_as(C.c_void_p)))

    @classmethod
    def from_sites(cls, ctx: Context, sites, lambdas=None, form: int = 0, dtype=None) -> "B200MPS":
        """Adapt an MPS given in the private layout: `sites[k]` has extents (chi_l, p, chi_r) (Fortran order,
        ideally pinned host memory), `lambdas[b]` the Schmidt vector on bond b or None."""
        n = len(sites)
        h = C.c_void_p()
        check(ctx.h, lib.qb200_mps_create_typed(ctx.h, n, _storage(sites, dtype), C.byref(h)))
        self = cls(ctx, _handle=h)
        for k, a in enumerate(sites):
            a, code = _host_code(a)
            assert a.ndim == 3
            a = np.asfortranarray(a)
            check(ctx.h, lib.qb200_mps_set_site_typed(ctx.h, h, k, code, a.shape[0], a.shape[1], a.shape[2],
                                                      a.ctypes.data_as(C.c_void_p)))
        for b, lam in enumerate(lambdas or []):
            if lam is not None:
                lam = np.ascontiguousarray(lam, dtype=np.float64)
                check(ctx.h, lib.qb200_mps_set_lambda(ctx.h, h, b, lam.shape[0],
                                                      lam.ctypes.data_as(C.POINTER(C.c_double))))
        check(ctx.h, lib.qb200_mps_set_form(h, form))
        return self

    @property
    def dtype(self):
        """Storage type of the site tensors in HBM."""
        return np.dtype(np.complex64) if lib.qb200_mps_dtype(self.h) == capi.C64 else np.dtype(np.complex128)

    @classmethod
    def from_product(cls, ctx: Context, vectors) -> "B200MPS":
        """`convert(Chain, ::Product)` (Chain.jl:174-183): a product state |v_1> ... |v_n> as a bond-dimension-1 chain.
        `a.overlap(b)` with it is the reference's `overlap(::Product, ::Chain)` (Chain.jl:751-752)."""
        vecs = [np.asarray(v, dtype=np.complex128).reshape(-1) for v in vectors]
        if len(vecs) < 2:
            raise ValueError("a chain needs at least two sites")
        sites = [np.asfortranarray(v.reshape(1, -1, 1)) for v in vecs]
        return cls.from_sites(ctx, sites)

    def site_into(self, s: int, out: np.ndarray):
        """Download site s into a caller-provided (pinned) Fortran-ordered buffer."""
        check(self.ctx.h, lib.qb200_mps_get_site(self.ctx.h, self.h, s, out.ctypes.data_as(C.c_void_p)))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib.qb200_mps_free(self.ctx.h, self.h)
        except Exception:
            pass
        self.h = None

    # -- bookkeeping -------------------------------------------------------------------------
    @property
    def nsites(self) -> int:
        return self._n()

    def _n(self):
        if not hasattr(self, "_nsites"):
            n = 0
            d = (C.c_int64 * 3)()
            while lib.qb200_mps_site_dims(self.h, n, d) == 0:
                n += 1
            self._nsites = n
        return self._nsites

    def copy(self) -> "B200MPS":
        h = C.c_void_p()
        check(self.ctx.h, lib.qb200_mps_copy(self.ctx.h, self.h, C.byref(h)))
        return B200MPS(self.ctx, _handle=h)

    def site_dims(self, s: int):
        d = (C.c_int64 * 3)()
        check(self.ctx.h, lib.qb200_mps_site_dims(self.h, s, d))
        return tuple(d)

    def bond_dims(self):
        return [self.site_dims(s)[2] for s in range(self.nsites - 1)]

    def site(self, s: int, dtype=np.complex128) -> np.ndarray:
        """Host copy of site s (0-based) with extents (chi_l, p, chi_r), as complex128 (default) or complex64."""
        d = self.site_dims(s)
        dt = np.dtype(dtype)
        out = np.empty(d, dtype=dt, order="F")
        check(self.ctx.h, lib.qb200_mps_get_site_typed(self.ctx.h, self.h, s, _HOST_CODES[dt],
                                                       out.ctypes.data_as(C.c_void_p)))
        return out

    def arrays(self):
        """Site arrays back in the reference's (o, l, r) order (edge bonds dropped)."""
        n = self.nsites
        out = []
        for s in range(n):
            a = np.transpose(self.site(s), (1, 0, 2))
            if s == 0:
                a = a[:, 0, :]
            if s == n - 1:
                a = a[..., 0]
            out.append(a)
        return out

    def lambdas(self):
        """Schmidt vector on each bond (None where absent) -- `tensors(ψ; between=(i, i+1))`."""
        out = []
        for b in range(self.nsites - 1):
            n = C.c_int64()
            code = lib.qb200_mps_get_lambda(self.ctx.h, self.h, b, None, C.byref(n))
            if code == capi.E_NOSPECTRUM:
                out.append(None)
                continue
            check(self.ctx.h, code)
            v = np.empty(n.value, dtype=np.float64)
            check(self.ctx.h, lib.qb200_mps_get_lambda(self.ctx.h, self.h, b, v.ctypes.data_as(C.POINTER(C.c_double)),
                                                       C.byref(n)))
            out.append(v)
        return out

    @property
    def form(self) -> int:
        return int(lib.qb200_mps_form(self.h))

    # -- Chain.jl algorithms -------------------------------------------------------------------
    def canonize(self) -> "B200MPS":
        """`canonize!` (Chain.jl:469-497)."""
        check(self.ctx.h, lib.qb200_mps_canonize(self.ctx.h, self.h))
        return self

    def mixed_canonize(self, center: int) -> "B200MPS":
        """`mixed_canonize!(ψ, Site(center))`, center 1-based as in the reference (Chain.jl:509-524)."""
        check(self.ctx.h, lib.qb200_mps_mixed_canonize(self.ctx.h, self.h, center - 1))
        return self

    def truncate(self, bond, threshold=None, maxdim=None) -> int:
        """`truncate!(ψ, [Site(i), Site(i+1)]; threshold, maxdim)`; bond = (i, i+1) 1-based (Chain.jl:390-422)."""
        i, j = bond
        if j != i + 1:
            raise ValueError(f"Invalid bond {bond}")
        kept = C.c_int64()
        check(self.ctx.h, lib.qb200_mps_truncate(self.ctx.h, self.h, i - 1, int(maxdim or 0),
                                                 -1.0 if threshold is None else float(threshold), C.byref(kept)))
        return kept.value

    def _lane_gate(self, gate, sites):
        """Gate array (dims (o_1.., i_1..), Dense.jl:21-34) for 1-based lanes `sites` -> (left site 0-based, nlanes,
        gate with its lanes in ascending order, flattened column-major)."""
        sites = list(sites)
        if len(sites) == 1:
            p = self.site_dims(sites[0] - 1)[1]
            g = np.asarray(gate, dtype=np.complex128).reshape((p, p), order="F")
            return sites[0] - 1, 1, np.reshape(g, -1, order="F")
        if len(sites) != 2:
            raise ValueError(f"Invalid number of lanes {len(sites)}, maximum is 2")
        a, b = sites
        if abs(a - b) != 1:
            raise ValueError("Gate lanes must be contiguous")
        pa, pb = self.site_dims(a - 1)[1], self.site_dims(b - 1)[1]
        g = np.asarray(gate, dtype=np.complex128).reshape((pa, pb, pa, pb), order="F")
        if a > b:  # gate given with lanes in descending order: swap its qubit roles
            g = np.transpose(g, (1, 0, 3, 2))
            a, b = b, a
        return a - 1, 2, np.reshape(g, -1, order="F")

    def evolve(self, gate, sites, threshold=None, maxdim=None, iscanonical=False, renormalize=False):
        """`evolve!(ψ, Dense(Operator(), gate; sites=[s.., s'..]); threshold, maxdim, iscanonical, renormalize)`
        (Chain.jl:543-584, same keyword defaults).  `gate` has array dims (o_1.., i_1..) as in the reference; `sites`
        are the 1-based lanes.  Returns (kept, discarded weight) for a 2-lane gate."""
        left, nl, flat = self._lane_gate(gate, sites)
        if nl == 1:
            check(self.ctx.h, lib.qb200_mps_evolve1(self.ctx.h, self.h, left, flat.ctypes.data_as(C.c_void_p)))
            return None
        kept = C.c_int64()
        dw = C.c_double()
        check(self.ctx.h, lib.qb200_mps_evolve2(self.ctx.h, self.h, left, flat.ctypes.data_as(C.c_void_p),
                                                int(maxdim or 0), -1.0 if threshold is None else float(threshold),
                                                int(bool(renormalize)), int(bool(iscanonical)), C.byref(kept),
                                                C.byref(dw)))
        return kept.value, dw.value

    def evolve_layer(self, gates, bonds, threshold=None, maxdim=None, iscanonical=False, renormalize=False):
        """One TEBD layer: `gates[i]` (dims (o1,o2,i1,i2)) on sites (bonds[i], bonds[i]+1), 1-based, pairwise
        non-adjacent -- the user loop `for b in bonds: evolve!(ψ, G_b; ...)` run as concurrent independent units.
        Returns (kept[], discarded_weight[])."""
        return self._evolve_ops(lib.qb200_mps_evolve2_layer, gates, bonds, threshold, maxdim, iscanonical, renormalize)

    def evolve_circuit(self, gates, bonds, threshold=None, maxdim=None, iscanonical=False, renormalize=False):
        """A gate list in program order: `gates[i]` on sites (bonds[i], bonds[i]+1), 1-based; bonds may repeat and
        touch -- the user loop `for (G, b) in circuit: evolve!(ψ, G; ...)` (Chain.jl:543-584).  Updates start as soon
        as the earlier updates on their two sites are done (consecutive layers overlap); results are identical to
        the sequential loop.  Returns (kept[], discarded_weight[])."""
        return self._evolve_ops(lib.qb200_mps_evolve2_circuit, gates, bonds, threshold, maxdim, iscanonical, renormalize)

    def _evolve_ops(self, fn, gates, bonds, threshold, maxdim, iscanonical, renormalize):
        nb = len(bonds)
        if len(gates) != nb:
            raise ValueError("one gate per bond expected")
        flat = np.concatenate([np.asarray(g, dtype=np.complex128).reshape(-1, order="F") for g in gates]) \
            if nb else np.zeros(0, np.complex128)
        kept = (C.c_int64 * max(nb, 1))()
        dw = (C.c_double * max(nb, 1))()
        check(self.ctx.h, fn(self.ctx.h, self.h, nb, capi.i32arr(b - 1 for b in bonds),
                             flat.ctypes.data_as(C.c_void_p), int(maxdim or 0),
                             -1.0 if threshold is None else float(threshold), int(bool(renormalize)),
                             int(bool(iscanonical)), kept, dw))
        return [int(kept[i]) for i in range(nb)], [float(dw[i]) for i in range(nb)]

    @staticmethod
    def _pack_mpo(mpo_arrays):
        """MPO arrays in the reference's default order (o, i, l, r) (first (o, i, r), last (o, i, l); Chain.jl:34,
        133-172) -> (dl, dr, flat column-major buffer)."""
        n = len(mpo_arrays)
        dl, dr, parts = [], [], []
        for k, w in enumerate(mpo_arrays):
            w = np.asarray(w, dtype=np.complex128)
            if k == 0:
                w = w[:, :, None, :] if w.ndim == 3 else w
            if k == n - 1:
                w = w[:, :, :, None] if w.ndim == 3 else w
            assert w.ndim == 4, (k, w.shape)
            dl.append(w.shape[2])
            dr.append(w.shape[3])
            parts.append(np.reshape(w, -1, order="F"))
        return capi.i64arr(dl), capi.i64arr(dr), np.ascontiguousarray(np.concatenate(parts))

    def apply_mpo(self, mpo_arrays) -> "B200MPS":
        """ψ <- H ψ site by site (`contract(merge(Quantum(ψ), Quantum(H)))` over the physical indices); bonds
        become χ·D.  Follow with `compress` to truncate."""
        dl, dr, flat = self._pack_mpo(mpo_arrays)
        check(self.ctx.h, lib.qb200_mps_apply_mpo(self.ctx.h, self.h, dl, dr, flat.ctypes.data_as(C.c_void_p)))
        return self

    def compress(self, maxdim=None, threshold=None) -> "B200MPS":
        """`canonize!` with `truncate!(…; maxdim, threshold)` applied to each bond right after its SVD."""
        check(self.ctx.h, lib.qb200_mps_compress(self.ctx.h, self.h, int(maxdim or 0),
                                                 -1.0 if threshold is None else float(threshold)))
        return self

    def expect_mpo(self, mpo_arrays) -> complex:
        """<ψ|H|ψ> = contract(merge(ψ, H, ψ')), un-normalised."""
        dl, dr, flat = self._pack_mpo(mpo_arrays)
        r = (C.c_double * 2)()
        check(self.ctx.h, lib.qb200_mps_expect_mpo(self.ctx.h, self.h, dl, dr, flat.ctypes.data_as(C.c_void_p), r))
        return complex(r[0], r[1])

    def overlap(self, other: "B200MPS") -> complex:
        """`overlap(a, b)` = <b|a> (Chain.jl:737-748)."""
        r = (C.c_double * 2)()
        check(self.ctx.h, lib.qb200_mps_overlap(self.ctx.h, self.h, other.h, r))
        return complex(r[0], r[1])

    def norm(self) -> float:
        """`norm(ψ)` (Ansatz.jl:101-109)."""
        return abs(np.sqrt(self.overlap(self)))

    def expect(self, ops, sites) -> np.ndarray:
        """Batch of single-site `expect(ψ, [O])` values, un-normalised (Chain.jl:724-735); sites 1-based.  All left
        and right environments are built once and shared by the batch."""
        ops = [np.asfortranarray(np.asarray(o, dtype=np.complex128)) for o in ops]
        flat = np.concatenate([o.reshape(-1, order="F") for o in ops]) if ops else np.zeros(0, np.complex128)
        res = np.zeros(2 * len(ops))
        check(self.ctx.h, lib.qb200_mps_expect1_batch(self.ctx.h, self.h, len(ops), capi.i32arr(s - 1 for s in sites),
                                                      flat.ctypes.data_as(C.c_void_p),
                                                      res.ctypes.data_as(C.POINTER(C.c_double))))
        return res[0::2] + 1j * res[1::2]

    def expect_observables(self, observables) -> complex:
        """`expect(ψ, observables)` for ANY list of 1- and 2-lane observables, with the reference's own composition
        (Chain.jl:724-735): ϕ = copy(ψ); evolve!(ϕ, O) for each O; contract(merge(ϕ, ψ')).  `observables` is a
        list of (array, lanes) with 1-based lanes, arrays as for `evolve`."""
        nl, left, parts = [], [], []
        for arr, lanes in observables:
            l, k, flat = self._lane_gate(arr, lanes)
            nl.append(k)
            left.append(l)
            parts.append(flat)
        flat = np.concatenate(parts) if parts else np.zeros(0, np.complex128)
        r = (C.c_double * 2)()
        check(self.ctx.h, lib.qb200_mps_expect(self.ctx.h, self.h, len(nl), capi.i32arr(nl), capi.i32arr(left),
                                               flat.ctypes.data_as(C.c_void_p), r))
        return complex(r[0], r[1])
