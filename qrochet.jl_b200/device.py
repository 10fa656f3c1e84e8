"""Device context and the device array type behind `Tenet.Tensor{T,N,B200Array}`: thin Python mirror of the
Julia extension's surface (julia/ext/QrochetB200Ext.jl) over the C-ABI.  Everything numeric happens inside
libqrochet_b200.so; NumPy is used only to stage host buffers."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import check, lib


class Context:
    """One context = one device + one stream (include/qrochet_b200.h conventions)."""

    def __init__(self, device: int = 0, stream=None):
        h = C.c_void_p()
        code = lib.qb200_create(device, C.byref(h))
        if code != 0:
            msg = lib.qb200_last_error(None)
            raise capi.QB200Error(code, (msg.decode() if msg else "") + " -- a B200 is required, no CPU fallback")
        self.h = h
        self.device = device
        if stream is not None:
            self.set_stream(stream)

    def set_stream(self, cuda_stream_ptr):
        check(self.h, lib.qb200_set_stream(self.h, C.c_void_p(int(cuda_stream_ptr) if cuda_stream_ptr else None)))

    def synchronize(self):
        check(self.h, lib.qb200_synchronize(self.h))

    @property
    def launches(self) -> int:
        return int(lib.qb200_launch_count(self.h))

    def timer_begin(self):
        check(self.h, lib.qb200_timer_begin(self.h))

    def timer_end(self) -> float:
        ms = C.c_double()
        check(self.h, lib.qb200_timer_end(self.h, C.byref(ms)))
        return ms.value

    PHASES = ("theta_gemm", "gate", "svd", "jacobi_gram", "jacobi_evd", "jacobi_update", "svd_emit", "mode_scale",
              "qr", "tn_gemm", "lowp_gram", "lowp_update", "lowp_glue")

    def profile(self, on: bool):
        check(self.h, lib.qb200_prof_enable(self.h, int(on)))

    def profile_read(self) -> dict:
        """{phase: (launches, total_ms, total_algorithmic_work)} since the last read."""
        n = len(self.PHASES)
        cnt, ms, work = (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        check(self.h, lib.qb200_prof_read(self.h, n, cnt, ms, work))
        return {name: (int(cnt[i]), float(ms[i]), float(work[i])) for i, name in enumerate(self.PHASES)}

    def dmma_peak_tflops(self) -> float:
        v = C.c_double()
        check(self.h, capi.diag().qb200_bench_dmma_peak(self.h, C.byref(v)))
        return v.value

    def hmma_peak_tflops(self):
        """(TF32, BF16) dense mma.sync peaks with FP32 accumulation -- denominators of the ComplexF32 kernels."""
        v = (C.c_double * 2)()
        check(self.h, capi.diag().qb200_bench_hmma_peak(self.h, v))
        return float(v[0]), float(v[1])

    def tcgen05_tf32_probe(self):
        """(max abs error of the self-checked tcgen05 TF32 product, TFLOP/s at N = 128, TFLOP/s at N = 256)."""
        v = (C.c_double * 3)()
        check(self.h, capi.diag().qb200_bench_tcgen05_tf32(self.h, v))
        return float(v[0]), float(v[1]), float(v[2])

    def tcgen05_i8_probe(self):
        """(max abs error of the self-checked tcgen05 INT8 product, TOP/s at N = 128, TOP/s at N = 256)."""
        v = (C.c_double * 3)()
        check(self.h, capi.diag().qb200_bench_tcgen05_i8(self.h, v))
        return float(v[0]), float(v[1]), float(v[2])

    def svd_totals(self):
        """(SVDs factorised, Jacobi sweeps executed) over this context and its worker streams since creation."""
        a, b = C.c_int64(), C.c_int64()
        check(self.h, lib.qb200_svd_totals(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def svd_last_sweeps(self) -> int:
        return int(lib.qb200_svd_last_sweeps(self.h))

    def close(self):
        if self.h:
            lib.qb200_destroy(self.h)
            self.h = None

    # -- array construction ------------------------------------------------------------------
    def empty(self, shape, dtype=capi.C128) -> "DeviceArray":
        return DeviceArray(self, tuple(int(s) for s in shape), dtype)

    def array(self, host) -> "DeviceArray":
        """`adapt(B200Array, ::Array)`: upload a host array (column-major semantics)."""
        host = np.asarray(host)
        if host.dtype == np.complex64:       # ComplexF32: native float2 storage, TF32-split tensor-core contraction
            buf, dt = np.asfortranarray(host), capi.C64
        elif host.dtype == np.float32:
            buf, dt = np.asfortranarray(host), capi.F32
        elif np.iscomplexobj(host):
            buf = np.asfortranarray(host, dtype=np.complex128)
            dt = capi.C128
        else:
            buf = np.asfortranarray(host, dtype=np.float64)
            dt = capi.F64
        if host.ndim == 0:  # np.asfortranarray promotes 0-d arrays to 1-d
            buf = buf.reshape(())
        out = DeviceArray(self, buf.shape, dt)
        if buf.size:
            check(self.h, lib.qb200_tensor_upload(self.h, out.h, buf.ctypes.data_as(C.c_void_p)))
        return out


class DeviceArray:
    """Dense column-major device array (ComplexF64, ComplexF32 or real) owned through a qb200_tensor handle."""

    def __init__(self, ctx: Context, shape, dtype=capi.C128, _handle=None):
        self.ctx = ctx
        self.dtype = dtype
        if _handle is None:
            h = C.c_void_p()
            check(ctx.h, lib.qb200_tensor_alloc(ctx.h, dtype, len(shape), capi.i64arr(shape), C.byref(h)))
            self.h = h
        else:
            self.h = _handle

    @property
    def shape(self):
        r = lib.qb200_tensor_rank(self.h)
        return tuple(int(lib.qb200_tensor_extent(self.h, i)) for i in range(r))

    @property
    def ndim(self):
        return int(lib.qb200_tensor_rank(self.h))

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                lib.qb200_tensor_free(self.ctx.h, self.h)
        except Exception:
            pass
        self.h = None

    def to_host(self) -> np.ndarray:
        shape = self.shape
        np_dtype = {capi.C128: np.complex128, capi.C64: np.complex64, capi.F64: np.float64, capi.F32: np.float32}
        out = np.empty(shape, dtype=np_dtype[self.dtype], order="F")
        if out.size:
            check(self.ctx.h, lib.qb200_tensor_download(self.ctx.h, self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def copy(self) -> "DeviceArray":
        out = DeviceArray(self.ctx, self.shape, self.dtype)
        check(self.ctx.h, lib.qb200_tensor_copy(self.ctx.h, self.h, out.h))
        return out

    def reshape(self, shape) -> "DeviceArray":
        """Column-major reshape of a copy-free view is not expressible with owning handles: returns a copy
        with the new shape (metadata-only on the copy)."""
        out = self.copy()
        check(self.ctx.h, lib.qb200_tensor_reshape(self.ctx.h, out.h, len(shape), capi.i64arr(shape)))
        return out

    def reshape_(self, shape) -> "DeviceArray":
        check(self.ctx.h, lib.qb200_tensor_reshape(self.ctx.h, self.h, len(shape), capi.i64arr(shape)))
        return self


# ---- functional surface (what the Julia extension dispatches to) ------------------------------------
def contract(a: DeviceArray, modes_a, b: DeviceArray, modes_b, modes_c, conj_a=False, conj_b=False,
             out: DeviceArray | None = None, alpha=1.0, beta=0.0) -> DeviceArray:
    ctx = a.ctx
    ext = {}
    for m, e in zip(modes_a, a.shape):
        ext[m] = e
    for m, e in zip(modes_b, b.shape):
        ext[m] = e
    if out is None:
        out = DeviceArray(ctx, tuple(ext[m] for m in modes_c), a.dtype)
    al = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
    be = (C.c_double * 2)(complex(beta).real, complex(beta).imag)
    check(ctx.h, lib.qb200_contract(ctx.h, a.h, capi.i32arr(modes_a), int(conj_a), b.h, capi.i32arr(modes_b),
                                    int(conj_b), out.h, capi.i32arr(modes_c), al, be))
    return out


def scale_mode(a: DeviceArray, mode_pos: int, vec: DeviceArray, inverse=False, atol=0.0, inplace=False) -> DeviceArray:
    out = a if inplace else DeviceArray(a.ctx, a.shape, a.dtype)
    check(a.ctx.h, lib.qb200_scale_mode(a.ctx.h, a.h, mode_pos, vec.h, int(inverse), float(atol), out.h))
    return out


def slice_mode(a: DeviceArray, mode_pos: int, count: int) -> DeviceArray:
    shape = list(a.shape)
    shape[mode_pos] = count
    out = DeviceArray(a.ctx, shape, a.dtype)
    check(a.ctx.h, lib.qb200_slice_mode(a.ctx.h, a.h, mode_pos, count, out.h))
    return out


def select_mode(a: DeviceArray, mode_pos: int, index: int) -> DeviceArray:
    shape = list(a.shape)
    del shape[mode_pos]
    out = DeviceArray(a.ctx, shape, a.dtype)
    check(a.ctx.h, lib.qb200_select_mode(a.ctx.h, a.h, mode_pos, index, out.h))
    return out


def conj(a: DeviceArray) -> DeviceArray:
    out = DeviceArray(a.ctx, a.shape, a.dtype)
    check(a.ctx.h, lib.qb200_conj(a.ctx.h, a.h, out.h))
    return out


def permute(a: DeviceArray, perm) -> DeviceArray:
    shape = a.shape
    out = DeviceArray(a.ctx, tuple(shape[p] for p in perm), a.dtype)
    check(a.ctx.h, lib.qb200_permute(a.ctx.h, a.h, capi.i32arr(perm), out.h))
    return out


def norm2(a: DeviceArray) -> float:
    v = C.c_double()
    check(a.ctx.h, lib.qb200_norm2(a.ctx.h, a.h, C.byref(v)))
    return v.value


def scale(a: DeviceArray, factor) -> DeviceArray:
    f = (C.c_double * 2)(complex(factor).real, complex(factor).imag)
    check(a.ctx.h, lib.qb200_scale(a.ctx.h, a.h, f))
    return a


def qr(a: DeviceArray, order, nleft: int):
    """Thin QR of the (left | right) matricisation; `order` lists mode positions, left modes first."""
    shape = a.shape
    rows = int(np.prod([shape[p] for p in order[:nleft]], dtype=np.int64))
    cols = int(np.prod([shape[p] for p in order[nleft:]], dtype=np.int64))
    k = min(rows, cols)
    q = DeviceArray(a.ctx, tuple(shape[p] for p in order[:nleft]) + (k,), a.dtype)
    r = DeviceArray(a.ctx, (k,) + tuple(shape[p] for p in order[nleft:]), a.dtype)
    check(a.ctx.h, lib.qb200_qr(a.ctx.h, a.h, capi.i32arr(order), nleft, q.h, r.h))
    return q, r


def svd(a: DeviceArray, order, nleft: int, maxdim: int = 0, threshold: float = -1.0):
    """Thin SVD with the truncate! rule applied; returns (U, S, Vc, kept, discarded_weight)."""
    shape = a.shape
    rows = int(np.prod([shape[p] for p in order[:nleft]], dtype=np.int64))
    cols = int(np.prod([shape[p] for p in order[nleft:]], dtype=np.int64))
    k = min(rows, cols)
    u = DeviceArray(a.ctx, tuple(shape[p] for p in order[:nleft]) + (k,), a.dtype)
    s = DeviceArray(a.ctx, (k,), capi.F32 if a.dtype == capi.C64 else capi.F64)  # Float32 spectrum for ComplexF32
    vc = DeviceArray(a.ctx, tuple(shape[p] for p in order[nleft:]) + (k,), a.dtype)
    kept = C.c_int64()
    dw = C.c_double()
    check(a.ctx.h, lib.qb200_svd(a.ctx.h, a.h, capi.i32arr(order), nleft, int(maxdim or 0), float(threshold), u.h, s.h,
                                 vc.h, C.byref(kept), C.byref(dw)))
    return u, s, vc, kept.value, dw.value
