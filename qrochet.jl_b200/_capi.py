"""ctypes binding of libqrochet_b200.so (include/qrochet_b200.h).  No CPU fallback: if the library or a
CUDA device is missing every call fails loudly."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libqrochet_b200.so")

C128, C64, F64, F32 = 0, 1, 2, 3
E_NOSPECTRUM = -4


class QB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libqrochet_b200 error {code}: {msg}")
        self.code = code


class MissingSchmidtCoefficientsException(QB200Error):
    """src/Ansatz.jl:91-99."""


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with ./build.sh (or __graft_entry__.build()); there is no CPU fallback"
        )
    return C.CDLL(LIB_PATH)


lib = _load()

_p = C.c_void_p
_i32, _i64, _dbl = C.c_int32, C.c_int64, C.c_double
_pi32, _pi64, _pdbl = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double)

_protos = {
    "qb200_create": (_i32, [_i32, C.POINTER(_p)]),
    "qb200_destroy": (_i32, [_p]),
    "qb200_last_error": (C.c_char_p, [_p]),
    "qb200_set_stream": (_i32, [_p, _p]),
    "qb200_synchronize": (_i32, [_p]),
    "qb200_launch_count": (_i64, [_p]),
    "qb200_timer_begin": (_i32, [_p]),
    "qb200_timer_end": (_i32, [_p, _pdbl]),
    "qb200_prof_enable": (_i32, [_p, _i32]),
    "qb200_prof_read": (_i32, [_p, _i32, _pi64, _pdbl, _pdbl]),
    "qb200_tensor_alloc": (_i32, [_p, _i32, _i32, _pi64, C.POINTER(_p)]),
    "qb200_tensor_wrap": (_i32, [_p, _i32, _i32, _pi64, _p, C.POINTER(_p)]),
    "qb200_tensor_free": (_i32, [_p, _p]),
    "qb200_tensor_upload": (_i32, [_p, _p, _p]),
    "qb200_tensor_download": (_i32, [_p, _p, _p]),
    "qb200_tensor_rank": (_i32, [_p]),
    "qb200_tensor_dtype": (_i32, [_p]),
    "qb200_tensor_extent": (_i64, [_p, _i32]),
    "qb200_tensor_data": (_p, [_p]),
    "qb200_tensor_copy": (_i32, [_p, _p, _p]),
    "qb200_tensor_reshape": (_i32, [_p, _p, _i32, _pi64]),
    "qb200_contract": (_i32, [_p, _p, _pi32, _i32, _p, _pi32, _i32, _p, _pi32, _pdbl, _pdbl]),
    "qb200_scale_mode": (_i32, [_p, _p, _i32, _p, _i32, _dbl, _p]),
    "qb200_slice_mode": (_i32, [_p, _p, _i32, _i64, _p]),
    "qb200_select_mode": (_i32, [_p, _p, _i32, _i64, _p]),
    "qb200_conj": (_i32, [_p, _p, _p]),
    "qb200_permute": (_i32, [_p, _p, _pi32, _p]),
    "qb200_norm2": (_i32, [_p, _p, _pdbl]),
    "qb200_scale": (_i32, [_p, _p, _pdbl]),
    "qb200_qr": (_i32, [_p, _p, _pi32, _i32, _p, _p]),
    "qb200_svd": (_i32, [_p, _p, _pi32, _i32, _i64, _dbl, _p, _p, _p, _pi64, _pdbl]),
    "qb200_svd_last_sweeps": (_i32, [_p]),
    "qb200_svd_totals": (_i32, [_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "qb200_mps_create": (_i32, [_p, _i32, C.POINTER(_p)]),
    "qb200_mps_create_typed": (_i32, [_p, _i32, _i32, C.POINTER(_p)]),
    "qb200_mps_dtype": (_i32, [_p]),
    "qb200_mps_set_site_typed": (_i32, [_p, _p, _i32, _i32, _i64, _i64, _i64, _p]),
    "qb200_mps_get_site_typed": (_i32, [_p, _p, _i32, _i32, _p]),
    "qb200_mps_free": (_i32, [_p, _p]),
    "qb200_mps_copy": (_i32, [_p, _p, C.POINTER(_p)]),
    "qb200_mps_set_site": (_i32, [_p, _p, _i32, _i64, _i64, _i64, _p]),
    "qb200_mps_site_dims": (_i32, [_p, _i32, _pi64]),
    "qb200_mps_get_site": (_i32, [_p, _p, _i32, _p]),
    "qb200_mps_set_lambda": (_i32, [_p, _p, _i32, _i64, _pdbl]),
    "qb200_mps_get_lambda": (_i32, [_p, _p, _i32, _pdbl, _pi64]),
    "qb200_mps_form": (_i32, [_p]),
    "qb200_mps_set_form": (_i32, [_p, _i32]),
    "qb200_mps_canonize": (_i32, [_p, _p]),
    "qb200_mps_mixed_canonize": (_i32, [_p, _p, _i32]),
    "qb200_mps_truncate": (_i32, [_p, _p, _i32, _i64, _dbl, _pi64]),
    "qb200_mps_evolve2": (_i32, [_p, _p, _i32, _p, _i64, _dbl, _i32, _i32, _pi64, _pdbl]),
    "qb200_mps_evolve2_layer": (_i32, [_p, _p, _i32, _pi32, _p, _i64, _dbl, _i32, _i32, _pi64, _pdbl]),
    "qb200_mps_evolve2_circuit": (_i32, [_p, _p, _i32, _pi32, _p, _i64, _dbl, _i32, _i32, _pi64, _pdbl]),
    "qb200_mps_evolve1": (_i32, [_p, _p, _i32, _p]),
    "qb200_mps_apply_mpo": (_i32, [_p, _p, _pi64, _pi64, _p]),
    "qb200_mps_compress": (_i32, [_p, _p, _i64, _dbl]),
    "qb200_mps_expect_mpo": (_i32, [_p, _p, _pi64, _pi64, _p, _pdbl]),
    "qb200_mps_overlap": (_i32, [_p, _p, _p, _pdbl]),
    "qb200_mps_expect1_batch": (_i32, [_p, _p, _i32, _pi32, _p, _pdbl]),
    "qb200_mps_expect": (_i32, [_p, _p, _i32, _pi32, _pi32, _p, _pdbl]),
    "qb200_tn_plan": (_i32, [_p, _i32, _pi32, _pi32, _pi64, _i64, C.POINTER(_p)]),
    "qb200_tn_plan_opt": (_i32, [_p, _i32, _pi32, _pi32, _pi64, _i64, _i32, C.POINTER(_p)]),
    "qb200_tn_plan_free": (_i32, [_p, _p]),
    "qb200_tn_plan_nslices": (_i64, [_p]),
    "qb200_tn_plan_sliced_modes": (_i32, [_p, _pi32]),
    "qb200_tn_plan_flops_per_slice": (_dbl, [_p]),
    "qb200_tn_plan_flops_invariant": (_dbl, [_p]),
    "qb200_tn_plan_max_intermediate": (_i64, [_p]),
    "qb200_tn_plan_path": (_i32, [_p, _pi32]),
    "qb200_tn_contract_sliced": (_i32, [_p, _p, C.POINTER(_p), _i64, _i64, _pdbl]),
    "qb200_comm_unique_id": (_i32, [C.c_char_p]),
    "qb200_comm_init": (_i32, [_p, _i32, _i32, C.c_char_p]),
    "qb200_comm_allreduce_sum": (_i32, [_p, _pdbl, _i32]),
    "qb200_comm_broadcast": (_i32, [_p, _p, _i32]),
    "qb200_mps_broadcast": (_i32, [_p, C.POINTER(_p), _i32]),
    "qb200_comm_destroy": (_i32, [_p]),
}

# micro-benchmarks / probes live in their own library (include/qrochet_b200_diag.h), loaded on first use
DIAG_LIB_PATH = os.path.join(_HERE, "lib", "libqrochet_b200_diag.so")
_diag_protos = {
    "qb200_bench_dmma_peak": (_i32, [_p, _pdbl]),
    "qb200_bench_hmma_peak": (_i32, [_p, _pdbl]),
    "qb200_bench_tcgen05_tf32": (_i32, [_p, _pdbl]),
    "qb200_bench_tcgen05_i8": (_i32, [_p, _pdbl]),
    "qb200_bench_dual_pipe": (_i32, [_p, _pdbl]),
    "qb200_bench_dmma_patterns": (_i32, [_p, _pdbl]),
    "qb200_bench_dmma_3m": (_i32, [_p, _pdbl]),
    "qb200_bench_update_variants": (_i32, [_p, _i32, _i32, _pdbl]),
}
DIAG_EXPORTS = sorted(_diag_protos)
_diag = None


def diag():
    """libqrochet_b200_diag.so (needs a CUDA device to do anything useful)."""
    global _diag
    if _diag is None:
        if not os.path.exists(DIAG_LIB_PATH):
            raise ImportError(f"{DIAG_LIB_PATH} not found: build it with ./build.sh")
        d = C.CDLL(DIAG_LIB_PATH)
        for name, (res, args) in _diag_protos.items():
            fn = getattr(d, name)
            fn.restype = res
            fn.argtypes = args
        _diag = d
    return _diag


EXPORTS = sorted(_protos)
for _name, (_res, _args) in _protos.items():
    _fn = getattr(lib, _name)  # AttributeError here = header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args


def check(ctx, code):
    if code != 0:
        msg = lib.qb200_last_error(ctx) if ctx is not None else b"(no context)"
        msg = msg.decode() if msg else ""
        if code == E_NOSPECTRUM:
            raise MissingSchmidtCoefficientsException(code, msg)
        raise QB200Error(code, msg)


def i32arr(xs):
    xs = list(xs)
    return (C.c_int32 * max(len(xs), 1))(*xs)


def i64arr(xs):
    xs = list(xs)
    return (C.c_int64 * max(len(xs), 1))(*xs)
