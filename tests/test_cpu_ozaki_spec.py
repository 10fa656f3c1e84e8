"""The bit-exact specification of the (experimental) INT8 panel product, tools/exp_ozaki.py::ozaki_complex: 8 slices of
7 bits reproduce the FP64 complex product to FP64 accuracy column by column, every slice fits a signed byte, and the
integer accumulators stay inside INT32 -- the properties csrc/i8_panel_gemm.cu relies on."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import exp_ozaki  # noqa: E402


def test_slices_fit_int8_and_reconstruct_exactly():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((40, 64)) * np.logspace(0, -12, 64)[None, :]
    a[3] = 0.0
    slices, e = exp_ozaki.split_rows(a, 8)
    assert all(np.abs(s).max() <= 64 for s in slices)
    rec = sum(np.ldexp(s.astype(np.float64), -7 * (t + 1)) for t, s in enumerate(slices)) * np.exp2(e)[:, None]
    assert np.abs(rec - a).max() <= 2.0 ** (-56) * np.abs(a).max(axis=1).max() * 2
    assert e[3] == 0 and not np.any(slices[0][3])


def test_integer_accumulators_stay_inside_int32():
    rng = np.random.default_rng(2)
    a, b = rng.standard_normal((128, 64)), rng.standard_normal((64, 64))
    acc, _, _ = exp_ozaki.ozaki_orders(a, b, 8)
    assert max(int(np.abs(x).max()) for x in acc) < 2 ** 31
    assert 8 * 64 * 64 * 64 < 2 ** 31          # worst case: 8 pairs of an order, K = 64, |slice| <= 64


def test_eight_slices_give_fp64_accuracy_column_wise():
    rng = np.random.default_rng(3)
    w = np.linalg.qr(rng.standard_normal((64, 64)) + 1j * rng.standard_normal((64, 64)))[0]
    for grade in (np.ones(64), np.logspace(0, -10, 64)):
        x = (rng.standard_normal((256, 64)) + 1j * rng.standard_normal((256, 64))) * grade[None, :]
        want = (x.astype(np.clongdouble) @ w.astype(np.clongdouble)).astype(np.complex128)
        assert exp_ozaki.colwise_err(exp_ozaki.ozaki_complex(x, w, 8), want) < 4e-15
