"""The opt-in WHILE-graph form of the Jacobi iteration (QB200_SVD_WHILE=1: one CUDA graph with a conditional node and
device-side convergence control) must take the same sweeps to the same factorisation.  The switch is read
once per process, so the check runs in a child."""
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
rng = np.random.default_rng(41)
for (m, n) in ((768, 512), (300, 520), (1024, 1024)):
    a = rng.standard_normal((m, n)) + 1j * rng.standard_normal((m, n))
    u, s, vc, kept, dw = qb.svd(ctx.array(np.asfortranarray(a)), (0, 1), 1)
    u, s, vc = u.to_host(), s.to_host(), vc.to_host()
    ref = sla.svd(a, compute_uv=False, lapack_driver="gesdd")
    k = min(m, n)
    assert kept == k and np.abs(s - ref).max() <= 1e-12 * ref[0]
    assert np.abs((u * s[None, :]) @ vc.T - a).max() <= 1e-11 * ref[0]
    print("ok", m, n, ctx.svd_last_sweeps())
"""


def test_while_graph_svd_matches_lapack_and_the_default_sweep_count():
    runs = {}
    for flag in ("1", "0"):
        env = dict(os.environ, QB200_SVD_WHILE=flag)
        out = subprocess.run([sys.executable, "-c", CODE % ROOT], capture_output=True, text=True, timeout=600, env=env)
        assert out.returncode == 0, out.stdout + out.stderr
        runs[flag] = [ln for ln in out.stdout.splitlines() if ln.startswith("ok")]
        assert len(runs[flag]) == 3
    assert runs["1"] == runs["0"]  # same sweep counts
