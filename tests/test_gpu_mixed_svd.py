"""The opt-in mixed-precision Jacobi SVD (QB200_SVD_MIXED=1: FP32 iteration on tcgen05 + FP64 finish) must
return the same factorisation as the FP64 path.  The switch is read once per process, so the check runs in a child."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CODE = r"""
import sys
import numpy as np
import scipy.linalg as sla
sys.path.insert(0, %r)
import qrochet_b200 as qb
ctx = qb.Context(0)
rng = np.random.default_rng(31)
for (m, n) in ((1024, 1024), (1536, 1024), (1024, 2048)):
    a = (rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))) * np.exp(-np.linspace(0, 6, n))[None, :]
    u, s, vc, kept, dw = qb.svd(ctx.array(np.asfortranarray(a)), (0, 1), 1)
    u, s, vc = u.to_host(), s.to_host(), vc.to_host()
    ref = sla.svd(a, compute_uv=False, lapack_driver="gesdd")
    k = min(m, n)
    assert kept == k, (kept, k)
    assert np.abs(s - ref).max() <= 1e-12 * ref[0], np.abs(s - ref).max() / ref[0]
    assert np.abs(u.conj().T @ u - np.eye(k)).max() <= 1e-10
    assert np.abs(vc.T @ vc.conj() - np.eye(k)).max() <= 1e-10
    assert np.abs((u * s[None, :]) @ vc.T - a).max() <= 1e-11 * ref[0]
    print("ok", m, n, ctx.svd_last_sweeps())
"""


@pytest.mark.parametrize("update", ["tc5", "simple"])
def test_mixed_precision_svd_matches_lapack(update):
    env = dict(os.environ, QB200_SVD_MIXED="1", QB200_C64_TCGEN05_CHECK="1")
    if update == "simple":
        env["QB200_LP_UPDATE"] = "simple"
    out = subprocess.run([sys.executable, "-c", CODE % ROOT], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 3
