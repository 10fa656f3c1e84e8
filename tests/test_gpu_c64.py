"""ComplexF32 path (BASELINE north star: "FP32/TF32-split tiles for ComplexF32", tolerance 1e-5 relative).

ComplexF32 tensors live on the device as float2; `contract` runs on tcgen05 (TF32, 3xTF32 split, TMEM accumulators
with two-level accumulation: gemm_c64_tc5.cu; the split-K shapes on the mma.sync kernel of gemm_c64.cu;
QB200_C64_TCGEN05=0 sends everything to the latter), the HBM-bound helpers run natively on float2, QR / SVD factorise
in FP64 and narrow the factors.
Everything is compared with FP64 NumPy on the same (float32-representable) inputs."""
import numpy as np
import pytest

from oracle import chain as oc

pytestmark = pytest.mark.gpu

TOL = 1e-5  # the north star's ComplexF32 tolerance


@pytest.fixture(scope="module")
def qb():
    import qrochet_b200 as q
    return q


@pytest.fixture(scope="module")
def ctx(qb):
    c = qb.Context(0)
    yield c
    c.close()


def crand32(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def test_storage_is_native_and_round_trips(qb, ctx):
    rng = np.random.default_rng(1)
    a = crand32(rng, 33, 17, 5)
    d = ctx.array(a)
    assert d.dtype == 1 and d.to_host().dtype == np.complex64
    assert np.array_equal(d.to_host(), a)                       # bit-exact: no widening round trip
    assert np.array_equal(d.copy().to_host(), a)
    v32 = ctx.array(rng.random(30).astype(np.float32))
    assert v32.to_host().dtype == np.float32


@pytest.mark.parametrize("shape", [(5, 7, 3), (130, 70, 66), (257, 300, 129), (1, 513, 1), (64, 8, 2048)])
@pytest.mark.parametrize("conj", [(False, False), (True, False), (False, True)])
def test_gemm_3xtf32_accuracy(qb, ctx, shape, conj):
    """Plain products incl. ragged tiles, split-K shapes and conj flags: FP32-level accuracy (a single-TF32 product
    would sit at ~1e-3)."""
    m, k, n = shape
    rng = np.random.default_rng(m * 7 + k)
    a, b = crand32(rng, m, k), crand32(rng, k, n)
    got = qb.contract(ctx.array(a), (0, 1), ctx.array(b), (1, 2), (0, 2), conj_a=conj[0], conj_b=conj[1]).to_host()
    assert got.dtype == np.complex64
    a64 = a.astype(np.complex128).conj() if conj[0] else a.astype(np.complex128)
    b64 = b.astype(np.complex128).conj() if conj[1] else b.astype(np.complex128)
    want = a64 @ b64
    # FP32-accumulation error model: ~ eps32 sqrt(K) |a| |b| per entry (plain TF32 would be ~ 1e-3 sqrt(K) |a| |b|)
    assert np.abs(got - want).max() <= 2e-6 * np.sqrt(k) * np.abs(a64).max() * np.abs(b64).max()
    assert np.abs(got - want).max() <= TOL * np.abs(want).max()


def test_gemm_large_k_beats_plain_tf32(qb, ctx):
    rng = np.random.default_rng(3)
    a, b = crand32(rng, 256, 4096), crand32(rng, 4096, 192)
    got = qb.contract(ctx.array(a), (0, 1), ctx.array(b), (1, 2), (0, 2)).to_host()
    want = a.astype(np.complex128) @ b.astype(np.complex128)
    rel = np.abs(got - want).max() / np.abs(want).max()
    assert rel < 5e-6, rel   # measured 1.9e-6 (tcgen05, two-level accumulation) / 2.8e-6 (mma.sync); a single TMEM
    # accumulator over K = 4096 gives 6.6e-5 (round-toward-zero accumulation) and plain TF32 ~3e-4


def test_alpha_beta_and_accumulate(qb, ctx):
    rng = np.random.default_rng(4)
    a, b, c = crand32(rng, 70, 40), crand32(rng, 40, 50), crand32(rng, 70, 50)
    out = ctx.array(c)
    qb.contract(ctx.array(a), (0, 1), ctx.array(b), (1, 2), (0, 2), out=out, alpha=0.5 - 2j, beta=1.5 + 0.25j)
    want = (0.5 - 2j) * (a.astype(np.complex128) @ b.astype(np.complex128)) + (1.5 + 0.25j) * c.astype(np.complex128)
    assert np.abs(out.to_host() - want).max() <= TOL * np.abs(want).max()


def test_contract_random_label_orders_c64(qb, ctx):
    """The seeded random einsum cases of the ComplexF64 suite (batch / summed / free modes, permuted labels, conj)."""
    rng = np.random.default_rng(2025)
    for case in range(40):
        n_m, n_n, n_k, n_b = rng.integers(0, 4), rng.integers(0, 4), rng.integers(0, 4), rng.integers(0, 2)
        labels = list(range(n_m + n_n + n_k + n_b))
        ext = {l: int(rng.integers(1, 6)) for l in labels}
        m_modes, rest = labels[:n_m], labels[n_m:]
        n_modes, rest = rest[:n_n], rest[n_n:]
        k_modes, b_modes = rest[:n_k], rest[n_k:]
        ma = [int(x) for x in rng.permutation(m_modes + k_modes + b_modes)]
        mb = [int(x) for x in rng.permutation(n_modes + k_modes + b_modes)]
        mc = [int(x) for x in rng.permutation(m_modes + n_modes + b_modes)]
        a = crand32(rng, *[ext[l] for l in ma]) if ma else np.array(crand32(rng, 1)[0])
        b = crand32(rng, *[ext[l] for l in mb]) if mb else np.array(crand32(rng, 1)[0])
        conj = (bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        a64 = a.astype(np.complex128).conj() if conj[0] else a.astype(np.complex128)
        b64 = b.astype(np.complex128).conj() if conj[1] else b.astype(np.complex128)
        want = np.einsum(a64, ma, b64, mb, mc)
        got = qb.contract(ctx.array(a), ma, ctx.array(b), mb, mc, conj_a=conj[0], conj_b=conj[1]).to_host()
        assert got.shape == np.shape(want), (case, ma, mb, mc)
        assert np.abs(got - want).max() <= TOL * max(1.0, np.abs(want).max()), (case, ma, mb, mc)


def test_mixed_types_are_rejected(qb, ctx):
    rng = np.random.default_rng(5)
    a = ctx.array(crand32(rng, 4, 4))
    b = ctx.array(crand32(rng, 4, 4).astype(np.complex128))
    with pytest.raises(Exception):
        qb.contract(a, (0, 1), b, (1, 2), (0, 2))


def test_hbm_helpers_on_float2(qb, ctx):
    """slice / select / conj / permute are bit-exact gathers; mode scale (with pinv) and norms apply FP64 factors."""
    rng = np.random.default_rng(6)
    a = crand32(rng, 6, 9, 4)
    d = ctx.array(a)
    assert np.array_equal(qb.slice_mode(d, 1, 5).to_host(), a[:, :5, :])
    assert np.array_equal(qb.select_mode(d, 2, 3).to_host(), a[:, :, 3])
    assert np.array_equal(qb.conj(d).to_host(), a.conj())
    assert np.array_equal(qb.permute(d, (2, 0, 1)).to_host(), np.transpose(a, (2, 0, 1)))
    v = rng.random(9) + 0.5
    v[2] = 1e-40
    want = a.astype(np.complex128) * v[None, :, None]
    got = qb.scale_mode(d, 1, ctx.array(v)).to_host()
    assert got.dtype == np.complex64 and np.abs(got - want).max() <= 1e-7 * np.abs(want).max()
    inv = np.where(np.abs(v) > 1e-32, 1.0 / v, 0.0)
    got = qb.scale_mode(d, 1, ctx.array(v.astype(np.float32)), inverse=True, atol=1e-32).to_host()
    want = a.astype(np.complex128) * inv[None, :, None]
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert abs(qb.norm2(d) - np.linalg.norm(a.astype(np.complex128))) <= 1e-12 * np.linalg.norm(a)
    s = qb.scale(d.copy(), 0.25 - 0.5j).to_host()
    assert np.abs(s - (0.25 - 0.5j) * a.astype(np.complex128)).max() <= 1e-6 * np.abs(a).max()


@pytest.mark.parametrize("shape,order,nleft", [((40, 30), (0, 1), 1), ((30, 40), (0, 1), 1),
                                               ((6, 5, 7), (2, 0, 1), 2), ((130, 64), (0, 1), 1)])
def test_svd_qr_of_c64_tensors(qb, ctx, shape, order, nleft):
    rng = np.random.default_rng(8)
    a = crand32(rng, *shape)
    mat = np.transpose(a, order).reshape(int(np.prod([shape[p] for p in order[:nleft]])), -1, order="F").astype(np.complex128)
    u, s, vc, kept, dw = qb.svd(ctx.array(a), order, nleft)
    s_ref = np.linalg.svd(mat, compute_uv=False)
    uh, sh, vh = u.to_host(), s.to_host(), vc.to_host()
    assert uh.dtype == np.complex64 and sh.dtype == np.float32 and vh.dtype == np.complex64
    k = len(s_ref)
    assert kept == k and np.abs(sh - s_ref).max() <= TOL * s_ref[0]
    um, vm = uh.reshape(-1, k, order="F").astype(np.complex128), vh.reshape(-1, k, order="F").astype(np.complex128)
    assert np.abs((um * sh.astype(np.float64)) @ vm.T - mat).max() <= TOL * s_ref[0]
    assert np.abs(um.conj().T @ um - np.eye(k)).max() <= TOL
    q, r = qb.qr(ctx.array(a), order, nleft)
    qm = q.to_host().reshape(-1, k, order="F").astype(np.complex128)
    rm = r.to_host().reshape(k, -1, order="F").astype(np.complex128)
    assert q.to_host().dtype == np.complex64
    assert np.abs(qm @ rm - mat).max() <= TOL * np.abs(mat).max() * np.sqrt(mat.shape[1])
    assert np.abs(qm.conj().T @ qm - np.eye(k)).max() <= TOL
    # truncation rule applied on the FP64 spectrum of the widened input
    u2, s2, v2, kept2, dw2 = qb.svd(ctx.array(a), order, nleft, maxdim=3)
    assert kept2 == min(3, k) and abs(dw2 - np.sum(s_ref[kept2:] ** 2)) <= TOL * np.sum(s_ref ** 2)


def test_label_driven_chain_in_complexf32(qb, ctx):
    """config 1 shape (n = 16, chi = 32) in ComplexF32 through the label-driven Chain: canonize!, overlap, <Z_8>,
    one evolve! -- against the FP64 oracle on the same inputs, 1e-5 relative."""
    ch, site = qb.chain, qb.chain.site
    n, chi = 16, 32
    arrays = [a.astype(np.complex64) for a in oc.rand_mps_arrays(np.random.default_rng(1001), n, chi)]
    arrays2 = [a.astype(np.complex64) for a in oc.rand_mps_arrays(np.random.default_rng(1002), n, chi)]
    q, q2 = ch.Chain(ctx, arrays), ch.Chain(ctx, arrays2)
    o = oc.Chain([a.astype(np.complex128) for a in arrays])
    o2 = oc.Chain([a.astype(np.complex128) for a in arrays2])
    assert q.eltype == np.complex64
    assert all(t.data.dtype in (1, 3) for t in q.tn.tensors)
    ov, ov_ref = q.overlap(q2), o.overlap(o2)
    assert abs(ov - ov_ref) <= TOL                              # relative to |a||b| = 1
    assert abs(q.norm() - o.norm()) <= TOL
    Z = np.diag([1.0, -1.0]).astype(complex)
    e_ref = o.expect([oc.gate(Z, [8])])
    assert abs(q.expect([(Z, [8])]) - e_ref) <= TOL                # relative to |psi|^2 |Z| = 1
    U = oc.haar_unitary(np.random.default_rng(9))
    G = np.reshape(U, (2, 2, 2, 2), order="F")
    q.evolve(G, [8, 9], maxdim=chi)
    o.evolve(oc.gate(U, [8, 9]), maxdim=chi)
    lam, lam_ref = q.lambda_between(site(8), site(9)).to_host(), o.lambdas()[7]
    kk = min(len(lam), len(lam_ref))
    assert np.abs(lam[:kk] - lam_ref[:kk]).max() <= TOL * lam_ref[0]
    assert all(t.data.dtype in (1, 3) for t in q.tn.tensors)


# ---- the fused chains (qb200_mps_*) with ComplexF32 storage and real element types (VERDICT r1 "missing" 4) ------------
def _mps_dense(g):
    lams, n = g.lambdas(), g.nsites
    psi = np.ones((1, 1), dtype=complex)
    for s in range(n):
        a = g.site(s)
        if s < n - 1 and lams[s] is not None:
            a = a * lams[s][None, None, :]
        psi = np.tensordot(psi, a, axes=(1, 0)).reshape(-1, a.shape[2], order="F")
    return psi[:, 0]


def test_fused_chain_with_complexf32_storage(qb, ctx):
    """`rand(Chain, Open, State; eltype = ComplexF32)` (Chain.jl:226-227) through canonize! / evolve! / truncate! / overlap /
    expect on the fused path: sites live in HBM as float2 (half the footprint), the kernels run in FP64, results are
    rounded to ComplexF32 once per call.  Parity bar of the north star for ComplexF32: 1e-5."""
    from oracle import chain as oc
    n, chi = 10, 16
    arrays = [a.astype(np.complex64) for a in oc.rand_mps_arrays(np.random.default_rng(501), n, chi)]
    o = oc.Chain([a.astype(np.complex128) for a in arrays])
    g = qb.B200MPS(ctx, arrays)
    assert g.dtype == np.complex64 and g.site(3, dtype=np.complex64).dtype == np.complex64
    assert np.array_equal(g.site(3, dtype=np.complex64), np.transpose(arrays[3], (1, 0, 2)))     # stored as given, bit for bit
    assert abs(g.norm() - o.norm()) <= TOL
    o.canonize()
    g.canonize()
    assert g.form == 1 and g.dtype == np.complex64
    for x, y in zip(g.lambdas(), o.lambdas()):
        assert len(x) == len(y) and np.abs(x - y).max() <= TOL * y[0]
    rng = np.random.default_rng(502)
    for bond in (5, 4, 6, 5):
        U = oc.haar_unitary(rng)
        o.evolve(oc.gate(U, [bond, bond + 1]), iscanonical=True, maxdim=12, renormalize=True)
        kept, _ = g.evolve(np.reshape(U, (2, 2, 2, 2), order="F"), [bond, bond + 1], iscanonical=True, maxdim=12,
                           renormalize=True)
        assert kept == len(o.lambdas()[bond - 1])                                  # truncation decisions: bit-exact
    for x, y in zip(g.lambdas(), o.lambdas()):
        assert np.abs(x - y).max() <= TOL * y[0]
    assert np.abs(_mps_dense(g) - o.to_dense()).max() <= 10 * TOL
    Z = np.diag([1.0, -1.0]).astype(complex)
    assert abs(g.expect([Z], [4])[0] - o.expect([oc.gate(Z, [4])])) <= TOL
    other = [a.astype(np.complex64) for a in oc.rand_mps_arrays(np.random.default_rng(503), n, 8)]
    want = o.overlap(oc.Chain([a.astype(np.complex128) for a in other]))
    assert abs(g.overlap(qb.B200MPS(ctx, other)) - want) <= TOL
    assert abs(g.overlap(qb.B200MPS(ctx, other, dtype=np.complex128)) - want) <= TOL     # mixed storage types
    # a gate list on worker streams and a copy keep the storage type
    gates = [np.reshape(oc.haar_unitary(rng), (2, 2, 2, 2), order="F") for _ in range(4)]
    c = g.copy()
    assert c.dtype == np.complex64
    c.evolve_circuit(gates, [2, 7, 3, 6], iscanonical=True, maxdim=12, renormalize=True)
    for gt, b in zip(gates, [2, 7, 3, 6]):
        o.evolve(oc.Dense(gt, [oc.site(b), oc.site(b + 1), oc.site(b, True), oc.site(b + 1, True)]), iscanonical=True,
                 maxdim=12, renormalize=True)
    for x, y in zip(c.lambdas(), o.lambdas()):
        assert len(x) == len(y) and np.abs(x - y).max() <= TOL * y[0]


def test_real_element_types_are_accepted_at_the_boundary(qb, ctx):
    """The reference's DEFAULT `eltype` of `rand` is Float64 (Chain.jl:226): real site arrays cross the boundary as they
    are and are held as complex."""
    from oracle import chain as oc
    n, chi = 8, 8
    for dt, tol in ((np.float64, 1e-10), (np.float32, TOL)):
        arrays = [a.astype(dt) for a in oc.rand_mps_arrays(np.random.default_rng(511), n, chi, dtype=np.float64)]
        o = oc.Chain([a.astype(np.complex128) for a in arrays])
        g = qb.B200MPS(ctx, arrays)
        assert g.dtype == (np.complex64 if dt == np.float32 else np.complex128)
        assert abs(g.norm() - o.norm()) <= tol
        o.canonize()
        g.canonize()
        for x, y in zip(g.lambdas(), o.lambdas()):
            assert np.abs(x - y).max() <= max(tol, 1e-12) * y[0]
        assert np.abs(np.abs(np.vdot(_mps_dense(g), o.to_dense())) - 1.0) <= 10 * tol
