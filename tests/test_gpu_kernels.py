"""GPU parity tests (through the C-ABI) of the individual kernels against NumPy/SciPy restatements:
K1 contraction, K2/K3 mode scale, K6 slice, K7 conj/permute, K8 norm, K4 QR, K5 SVD."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qb():
    import qrochet_b200 as q
    return q


@pytest.fixture(scope="module")
def ctx(qb):
    c = qb.Context(0)
    yield c
    c.close()


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def einsum_ref(a, ma, b, mb, mc, conj_a=False, conj_b=False):
    if conj_a:
        a = a.conj()
    if conj_b:
        b = b.conj()
    return np.einsum(a, list(ma), b, list(mb), list(mc))


CONTRACT_CASES = [
    # (shape_a, modes_a, shape_b, modes_b, modes_c)
    ((5, 7), (0, 1), (7, 3), (1, 2), (0, 2)),                      # plain matmul, ragged
    ((130, 70), (0, 1), (70, 66), (1, 2), (0, 2)),                 # crosses tile boundaries
    ((70, 130), (1, 0), (66, 70), (2, 1), (2, 0)),                 # both transposed, C transposed
    ((4, 6, 5), (0, 1, 2), (5, 3, 6), (2, 3, 1), (3, 0)),          # two summed modes, permuted
    ((2, 3, 4, 5), (0, 1, 2, 3), (4, 2, 6), (2, 0, 6), (3, 6, 1)),  # interleaved M/K modes
    ((3, 4, 5), (0, 1, 2), (4, 6, 3), (1, 6, 0), (0, 2, 6)),       # batch (kept shared) mode 0
    ((8,), (0,), (8,), (0,), ()),                                  # inner product -> scalar
    ((6,), (0,), (7,), (1,), (1, 0)),                              # outer product (K = 1)
    ((2,) * 10, tuple(range(10)), (2,) * 8, (1, 3, 5, 7, 9, 10, 11, 12), (0, 12, 2, 11, 4, 10, 6, 8)),  # circuit-like
    ((1, 300000), (0, 1), (300000,), (1,), (0,)),                  # long K: split-K path
    ((300, 20), (0, 1), (20, 200), (1, 2), (0, 2)),                # M > N
    ((20, 200), (1, 0), (20, 300), (1, 2), (0, 2)),                # N > M -> swapped operands
    # gate-sized operand against a long free dimension: the thin kernel (N <= 16, M >= 2048 after orientation)
    ((4100, 8), (0, 1), (8, 8), (1, 2), (0, 2)),                   # N = 8, K = 8, ragged M
    ((8, 4100), (1, 0), (5, 8), (2, 1), (2, 0)),                   # transposed operands and result, N = 5
    ((2,) * 13, tuple(range(13)), (2,) * 6, (1, 5, 9, 20, 21, 22),
     (0, 20, 2, 3, 4, 21, 6, 7, 8, 22, 10, 11, 12)),               # circuit-like: 3-qubit gate on a 13-qubit state
    ((2,) * 14, tuple(range(14)), (4, 4, 2, 2), (30, 31, 3, 11),
     (0, 1, 2, 30, 4, 5, 6, 7, 8, 9, 10, 31, 12, 13)),             # N = 16 (two 8-column tiles), K = 4
    ((16, 130, 3), (0, 1, 2), (130, 12), (1, 5), (5, 0, 2)),        # K = 130 (17 k tiles), batchless, N = 12
    ((2100, 3, 20), (0, 1, 2), (20, 3, 7), (2, 1, 5), (1, 0, 5)),   # kept shared mode = batch of thin products
    ((2,) * 14, tuple(range(14)), (2,) * 9, (2, 6, 10, 20, 21, 22, 23, 24, 25),
     (0, 1, 20, 3, 4, 5, 21, 7, 8, 9, 22, 11, 12, 13, 23, 24, 25)),  # N = 64 in four 16-column tiles, K = 8
    ((3, 5000, 2), (0, 1, 2), (4, 2, 5000), (3, 2, 1), (3, 0)),     # <= 4 x 4 outputs over K = 10000: the dot kernel
    ((2,) * 14, tuple(range(14)), (2,) * 14, (0, 20, 2, 3, 4, 5, 6, 7, 21, 9, 10, 11, 12, 13), (8, 1, 20, 21)),  # closing contraction
]


@pytest.mark.parametrize("case", range(len(CONTRACT_CASES)))
@pytest.mark.parametrize("conj", [(False, False), (True, False), (False, True)])
def test_contract_matches_einsum(qb, ctx, case, conj):
    sa, ma, sb, mb, mc = CONTRACT_CASES[case]
    rng = np.random.default_rng(100 + case)
    a, b = crand(rng, *sa), crand(rng, *sb)
    want = einsum_ref(a, ma, b, mb, mc, *conj)
    got = qb.contract(ctx.array(a), ma, ctx.array(b), mb, mc, conj_a=conj[0], conj_b=conj[1]).to_host()
    scale = max(1.0, np.abs(want).max())
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-12 * scale * max(1, int(np.prod(sa)) ** 0.5 / 30)


def test_contract_alpha_beta(qb, ctx):
    rng = np.random.default_rng(1)
    a, b, c = crand(rng, 40, 30), crand(rng, 30, 50), crand(rng, 40, 50)
    out = ctx.array(c)
    qb.contract(ctx.array(a), (0, 1), ctx.array(b), (1, 2), (0, 2), out=out, alpha=2 - 1j, beta=0.5j)
    assert np.allclose(out.to_host(), (2 - 1j) * a @ b + 0.5j * c, atol=1e-11)


@pytest.mark.parametrize("shape", [(4100, 8, 8), (2300, 12, 40), (2100, 7, 130), (3, 4, 6000), (2500, 64, 16), (2500, 64, 64)])
def test_contract_alpha_beta_on_the_shape_specialised_kernels(qb, ctx, shape):
    """alpha / beta epilogues of the streaming (N <= 16, K <= 32), thin (N <= 16 longer K; N <= 64 at K <= 16), dot (<= 4 x 4
    outputs, long K) and 64-row tile kernels."""
    m, n, k = shape
    rng = np.random.default_rng(sum(shape))
    a, b, c = crand(rng, m, k), crand(rng, k, n), crand(rng, m, n)
    out = ctx.array(c)
    qb.contract(ctx.array(a), (0, 1), ctx.array(b), (1, 2), (0, 2), out=out, alpha=1.5 - 0.5j, beta=-0.25 + 2j)
    want = (1.5 - 0.5j) * a @ b + (-0.25 + 2j) * c
    assert np.abs(out.to_host() - want).max() <= 1e-12 * np.abs(want).max() * max(1.0, k ** 0.5)
    out1 = ctx.array(c)
    qb.contract(ctx.array(a), (0, 1), ctx.array(b), (1, 2), (0, 2), out=out1, alpha=1.0, beta=1.0)
    assert np.abs(out1.to_host() - (a @ b + c)).max() <= 1e-12 * np.abs(want).max() * max(1.0, k ** 0.5)


@pytest.mark.parametrize("alpha", [1.0, -1.0])
def test_contract_accumulate_in_place(qb, ctx, alpha):
    """C += (+-1) A B (beta = 1): the accumulators start from +-C (the QR trailing update T -= P C)."""
    rng = np.random.default_rng(11)
    a, b, c = crand(rng, 200, 64), crand(rng, 64, 150), crand(rng, 200, 150)
    out = ctx.array(c)
    qb.contract(ctx.array(a), (0, 1), ctx.array(b), (1, 2), (0, 2), out=out, alpha=alpha, beta=1.0)
    assert np.allclose(out.to_host(), c + alpha * (a @ b), atol=1e-11)


def test_contract_errors(qb, ctx):
    a, b = ctx.array(np.ones((2, 3), complex)), ctx.array(np.ones((4, 5), complex))
    with pytest.raises(qb.QB200Error):
        qb.contract(a, (0, 1), b, (1, 2), (0, 2))  # extent mismatch on mode 1
    with pytest.raises(qb.QB200Error):
        qb.contract(a, (0, 0), b, (1, 2), (0, 2))  # repeated mode


def test_elementwise(qb, ctx):
    rng = np.random.default_rng(2)
    a = crand(rng, 6, 5, 7)
    v = rng.random(5)
    v[2] = 1e-40
    d, dv = ctx.array(a), ctx.array(v)
    assert np.allclose(qb.scale_mode(d, 1, dv).to_host(), a * v[None, :, None])
    inv = np.where(np.abs(v) > 1e-32, 1 / v, 0.0)
    assert np.allclose(qb.scale_mode(d, 1, dv, inverse=True, atol=1e-32).to_host(), a * inv[None, :, None])
    assert np.array_equal(qb.slice_mode(d, 2, 3).to_host(), a[:, :, :3])
    assert np.array_equal(qb.slice_mode(d, 0, 6).to_host(), a)
    assert np.array_equal(qb.select_mode(d, 1, 4).to_host(), a[:, 4, :])
    assert np.array_equal(qb.conj(d).to_host(), a.conj())
    assert np.array_equal(qb.permute(d, (2, 0, 1)).to_host(), np.transpose(a, (2, 0, 1)))
    assert np.isclose(qb.norm2(d), np.linalg.norm(a), rtol=1e-14)
    assert np.allclose(qb.scale(d.copy(), 0.5 - 2j).to_host(), a * (0.5 - 2j))
    assert np.array_equal(d.reshape((30, 7)).to_host(), a.reshape((30, 7), order="F"))
    empty = qb.slice_mode(d, 1, 0)
    assert empty.shape == (6, 0, 7)


QR_SHAPES = [(4, 4), (64, 32), (33, 7), (200, 64), (500, 130), (130, 500), (2, 4), (1000, 96)]


@pytest.mark.parametrize("shape", QR_SHAPES)
def test_qr(qb, ctx, shape):
    rng = np.random.default_rng(3)
    a = crand(rng, *shape)
    q, r = qb.qr(ctx.array(a), (0, 1), 1)
    q, r = q.to_host(), r.to_host()
    k = min(shape)
    assert q.shape == (shape[0], k) and r.shape == (k, shape[1])
    assert np.abs(q.conj().T @ q - np.eye(k)).max() < 1e-13
    assert np.abs(q @ r - a).max() < 1e-12 * np.abs(a).max() * 10
    assert np.abs(np.tril(r[:, :k], -1)).max() < 1e-13
    # R is unique up to row phases: compare |R| with LAPACK (gauge-invariant)
    r_ref = sla.qr(a, mode="economic")[1]
    assert np.allclose(np.abs(r), np.abs(r_ref), atol=1e-11 * np.abs(a).max() * 10)


def test_qr_rank_deficient_and_permuted(qb, ctx):
    rng = np.random.default_rng(4)
    base = crand(rng, 96, 10)
    a = base @ crand(rng, 10, 40)  # rank 10
    q, r = qb.qr(ctx.array(a), (0, 1), 1)
    q, r = q.to_host(), r.to_host()
    assert np.abs(q.conj().T @ q - np.eye(40)).max() < 1e-12
    assert np.abs(q @ r - a).max() < 1e-11
    t = crand(rng, 3, 4, 5)
    q, r = qb.qr(ctx.array(t), (2, 0, 1), 2)  # left = (mode2, mode0), right = mode1
    mat = np.transpose(t, (2, 0, 1)).reshape((15, 4), order="F")
    assert np.allclose(q.to_host().reshape((15, 4), order="F") @ r.to_host(), mat)


SVD_SHAPES = [(4, 4), (2, 8), (64, 32), (33, 70), (100, 130), (256, 256), (300, 129), (512, 512)]


@pytest.mark.parametrize("shape", SVD_SHAPES)
def test_svd(qb, ctx, shape):
    rng = np.random.default_rng(5)
    a = crand(rng, *shape)
    u, s, vc, kept, dw = qb.svd(ctx.array(a), (0, 1), 1)
    u, s, vc = u.to_host(), s.to_host(), vc.to_host()
    k = min(shape)
    s_ref = sla.svd(a, compute_uv=False, lapack_driver="gesdd")
    assert kept == k and s.shape == (k,) and dw == 0.0
    assert np.abs(s - s_ref).max() <= 1e-12 * s_ref[0]          # north-star tolerance on sigma
    assert np.all(np.diff(s) <= 0)
    assert np.abs(u.conj().T @ u - np.eye(k)).max() < 1e-11
    assert np.abs(vc.T @ vc.conj() - np.eye(k)).max() < 1e-11
    assert np.abs((u * s) @ vc.T - a).max() < 1e-12 * s_ref[0]


def test_svd_graded_spectrum_and_truncation(qb, ctx):
    rng = np.random.default_rng(6)
    n = 96
    q1, _ = np.linalg.qr(crand(rng, n, n))
    q2, _ = np.linalg.qr(crand(rng, n, n))
    sig = np.logspace(0, -14, n)
    a = (q1 * sig) @ q2.conj().T
    u, s, vc, kept, dw = qb.svd(ctx.array(a), (0, 1), 1, maxdim=20)
    s_ref = sla.svd(a, compute_uv=False, lapack_driver="gesdd")
    assert kept == 20 and s.shape == (20,) and u.shape == (n, 20) and vc.shape == (n, 20)
    assert np.abs(s.to_host() - s_ref[:20]).max() <= 1e-12 * s_ref[0]
    assert np.isclose(dw, np.sum(s_ref[20:] ** 2), rtol=1e-9)
    # threshold rule of truncate! (Chain.jl:411-417): s[i] > threshold, bit-exact count
    thr = float(s_ref[9] * 0.999)
    _, s2, _, kept2, _ = qb.svd(ctx.array(a), (0, 1), 1, threshold=thr)
    assert kept2 == int(np.sum(s_ref > thr)) == 10
    # both limits apply
    _, _, _, kept3, _ = qb.svd(ctx.array(a), (0, 1), 1, maxdim=4, threshold=thr)
    assert kept3 == 4


def test_svd_rank_deficient_and_tensor_modes(qb, ctx):
    rng = np.random.default_rng(7)
    a = crand(rng, 80, 6) @ crand(rng, 6, 50)
    u, s, vc, kept, dw = qb.svd(ctx.array(a), (0, 1), 1, threshold=1e-10)
    s_ref = sla.svd(a, compute_uv=False)
    assert kept == 6
    assert np.abs(s.to_host() - s_ref[:6]).max() <= 1e-12 * s_ref[0]
    assert np.abs((u.to_host() * s.to_host()) @ vc.to_host().T - a).max() < 1e-11 * s_ref[0]
    t = crand(rng, 3, 4, 5)
    u, s, vc, kept, _ = qb.svd(ctx.array(t), (1, 2, 0), 1)  # left = mode1 ; right = (mode2, mode0)
    mat = np.transpose(t, (1, 2, 0)).reshape((4, 15), order="F")
    rec = (u.to_host() * s.to_host()) @ vc.to_host().reshape((15, 4), order="F").T
    assert np.allclose(rec, mat)


def test_svd_graded_kept_spectrum_keeps_factors_orthonormal(qb, ctx):
    """Kept singular values spanning 10 orders of magnitude: U and V stay orthonormal, A = U S V^H holds norm-wise,
    sigma matches LAPACK to 1e-12 sigma_1 (the rotations are not accumulated; V is recovered from R X S^-1 and
    re-orthonormalised when the kept spectrum is graded)."""
    rng = np.random.default_rng(8)
    for (m, n) in [(160, 160), (200, 96), (96, 200)]:
        k = min(m, n)
        q1, _ = np.linalg.qr(crand(rng, m, k))
        q2, _ = np.linalg.qr(crand(rng, n, k))
        sig = np.logspace(0, -10, k)
        a = (q1 * sig) @ q2.conj().T
        u, s, vc, kept, dw = qb.svd(ctx.array(a), (0, 1), 1)
        u, s, vc = u.to_host(), s.to_host(), vc.to_host()
        assert np.abs(s - sig).max() <= 1e-12
        assert np.abs(u.conj().T @ u - np.eye(k)).max() < 1e-10
        assert np.abs(vc.T @ vc.conj() - np.eye(k)).max() < 1e-6    # small-sigma vectors: eps sigma_1 / sigma_j
        assert np.abs((u * s) @ vc.T - a).max() < 1e-12


def test_complexf32_tensors_cross_the_boundary(qb, ctx):
    """ComplexF32 / Float32 arrays are accepted (widened to FP64 on the device, FP64 arithmetic): the reference's
    ComplexF32 tolerance is 1e-5 relative."""
    rng = np.random.default_rng(12)
    a = crand(rng, 40, 30).astype(np.complex64)
    b = crand(rng, 30, 20).astype(np.complex64)
    da, db = ctx.array(a), ctx.array(b)
    assert da.to_host().dtype == np.complex64 and np.array_equal(da.to_host(), a)
    c = qb.contract(da, (0, 1), db, (1, 2), (0, 2)).to_host()
    want = a.astype(np.complex128) @ b.astype(np.complex128)
    assert np.abs(c - want).max() <= 1e-5 * np.abs(want).max()
    u, s, vc, kept, _ = qb.svd(da, (0, 1), 1)
    s_ref = np.linalg.svd(a.astype(np.complex128), compute_uv=False)
    assert np.abs(s.to_host() - s_ref).max() <= 1e-5 * s_ref[0]
    v32 = ctx.array(rng.random(30).astype(np.float32))
    assert v32.to_host().dtype == np.float32


def test_contract_random_label_orders(qb, ctx):
    """Seeded random einsum cases: random ranks, extents, label permutations, batch / summed / free modes and conj
    flags (what a Tenet contraction path throws at K1), each checked against numpy.einsum."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        n_m, n_n, n_k, n_b = rng.integers(0, 4), rng.integers(0, 4), rng.integers(0, 4), rng.integers(0, 2)
        labels = list(range(n_m + n_n + n_k + n_b))
        ext = {l: int(rng.integers(1, 6)) for l in labels}
        m_modes, rest = labels[:n_m], labels[n_m:]
        n_modes, rest = rest[:n_n], rest[n_n:]
        k_modes, b_modes = rest[:n_k], rest[n_k:]
        ma = list(rng.permutation(m_modes + k_modes + b_modes))
        mb = list(rng.permutation(n_modes + k_modes + b_modes))
        mc = list(rng.permutation(m_modes + n_modes + b_modes))
        a = crand(rng, *[ext[l] for l in ma]) if ma else np.array(crand(rng, 1)[0])
        b = crand(rng, *[ext[l] for l in mb]) if mb else np.array(crand(rng, 1)[0])
        conj = (bool(rng.integers(0, 2)), bool(rng.integers(0, 2)))
        want = einsum_ref(a, [int(x) for x in ma], b, [int(x) for x in mb], [int(x) for x in mc], *conj)
        got = qb.contract(ctx.array(a), [int(x) for x in ma], ctx.array(b), [int(x) for x in mb],
                          [int(x) for x in mc], conj_a=conj[0], conj_b=conj[1]).to_host()
        assert got.shape == np.shape(want), (case, ma, mb, mc)
        assert np.abs(got - want).max() <= 1e-12 * max(1.0, np.abs(want).max()), (case, ma, mb, mc)


@pytest.mark.parametrize("name", ["zero", "identity", "rank1", "skinny", "scalar", "row", "column", "repeated"])
def test_svd_degenerate_inputs(qb, ctx, name):
    rng = np.random.default_rng(13)
    a = {
        "zero": np.zeros((40, 24), complex),
        "identity": np.eye(70, dtype=complex),
        "rank1": np.outer(crand(rng, 90), crand(rng, 33)),
        "skinny": crand(rng, 1000, 3),
        "scalar": np.array([[2.0 - 1.0j]]),
        "row": crand(rng, 1, 50),
        "column": crand(rng, 50, 1),
        "repeated": np.kron(np.eye(4), crand(rng, 16, 16)),       # every singular value four-fold degenerate
    }[name]
    u, s, vc, kept, dw = qb.svd(ctx.array(a), (0, 1), 1)
    u, s, vc = u.to_host(), s.to_host(), vc.to_host()
    s_ref = np.linalg.svd(a, compute_uv=False)
    scale = max(s_ref[0], 1e-300)
    assert np.all(np.isfinite(u)) and np.all(np.isfinite(vc)) and np.all(np.isfinite(s))
    assert np.abs(s - s_ref).max() <= 1e-12 * scale
    assert np.abs((u * s) @ vc.T - a).max() <= 1e-12 * max(scale, 1.0)
    r = int(np.sum(s_ref > 1e-10 * scale))                     # the numerically non-zero part is orthonormal
    if r:
        assert np.abs(u[:, :r].conj().T @ u[:, :r] - np.eye(r)).max() < 1e-11
        assert np.abs(vc[:, :r].T @ vc[:, :r].conj() - np.eye(r)).max() < 1e-11


@pytest.mark.parametrize("shape", [(1, 1), (5, 1), (1, 5), (65, 64), (64, 65), (129, 3)])
def test_qr_edge_shapes(qb, ctx, shape):
    rng = np.random.default_rng(14)
    a = crand(rng, *shape)
    q, r = qb.qr(ctx.array(a), (0, 1), 1)
    q, r = q.to_host(), r.to_host()
    k = min(shape)
    assert np.abs(q.conj().T @ q - np.eye(k)).max() < 1e-13
    assert np.abs(q @ r - a).max() < 1e-12 * max(1.0, np.abs(a).max())


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64, np.float64])
def test_gather_paths_bit_exact(qb, ctx, dtype):
    """slice / select / conj / permute are pure data movement: bit-exact against NumPy on every code path of the
    gather -- merged modes, 16-byte pairs, 32-bit magic-number index decomposition, the 32 x 32 tiled transpose
    (full, ragged and many-mode cases) and extent-1 modes."""
    rng = np.random.default_rng(77)

    def rnd(*shape):
        a = rng.standard_normal(shape)
        if np.issubdtype(dtype, np.complexfloating):
            a = a + 1j * rng.standard_normal(shape)
        return a.astype(dtype)

    cases = [((33, 70), (1, 0)), ((64, 64), (1, 0)), ((40, 3, 50), (2, 1, 0)), ((40, 3, 50), (2, 0, 1)),
             ((9, 8, 7, 10), (3, 1, 0, 2)), ((16, 2, 16), (1, 0, 2)), ((5, 1, 12, 1, 9), (4, 3, 2, 1, 0)),
             ((2, 2, 2, 2, 2, 2, 2, 2, 2, 2), (9, 0, 8, 1, 7, 2, 6, 3, 5, 4)), ((100,), (0,)), ((8, 129), (1, 0))]
    for shape, perm in cases:
        a = rnd(*shape)
        d = ctx.array(a)
        assert np.array_equal(qb.permute(d, perm).to_host(), np.transpose(a, perm)), (shape, perm)
        if np.issubdtype(dtype, np.complexfloating):
            assert np.array_equal(qb.conj(d).to_host(), a.conj())
        for pos in range(len(shape)):
            cnt = max(1, shape[pos] // 2)
            idx = [slice(None)] * len(shape)
            idx[pos] = slice(0, cnt)
            assert np.array_equal(qb.slice_mode(d, pos, cnt).to_host(), a[tuple(idx)]), (shape, pos)
            idx[pos] = shape[pos] - 1
            assert np.array_equal(qb.select_mode(d, pos, shape[pos] - 1).to_host(), a[tuple(idx)]), (shape, pos)
    v = rnd(257)
    assert abs(qb.norm2(ctx.array(v)) - np.linalg.norm(v.astype(np.complex128 if np.iscomplexobj(v) else np.float64))) \
        <= 1e-12 * np.linalg.norm(v)


def test_tcgen05_tf32_building_block_is_exact(qb, ctx):
    """The hand-written tcgen05 / TMEM primitives under the ComplexF32 GEMM (tc5_probe.cu: UMMA shared-memory layout by
    st.shared, tcgen05.alloc, mma.kind::tf32 from one thread, commit -> mbarrier, tcgen05.ld): an integer-valued
    128 x N x 96 product must come back exactly, and the issue-bound rate must be in tcgen05 territory (the mma.sync
    TF32 path peaks at 277 TFLOP/s on this part)."""
    err, tf128, tf256 = ctx.tcgen05_tf32_probe()
    assert err == 0.0
    assert tf128 > 500.0 and tf256 > 500.0, (tf128, tf256)


def _structured_rank_deficient(rng, m, n, r):
    """Rank r with EXACT zeros and exact repetitions (the structure of an MPO-applied site tensor): what is left of a
    dependent column after projecting out the others is rounding noise that lies INSIDE the span already covered --
    the case in which normalising the noise (Gram-Schmidt, round 1) gave Q columns that were not orthogonal."""
    a = np.zeros((m, n), dtype=complex)
    half = m // 2
    base = crand(rng, half, r)
    a[:half, :r] = base
    for j in range(r, n):                      # every further column: an exact copy or an exact combination, rows below `half` all zero
        a[:half, j] = base[:, j % r] if j % 3 else base[:, j % r] + 2.0 * base[:, (j + 1) % r]
    return a


@pytest.mark.parametrize("m,n,r", [(1280, 640, 200), (640, 640, 256), (200, 120, 30), (2560, 1280, 700)])
def test_qr_of_structured_rank_deficient_input_is_orthonormal(qb, ctx, m, n, r):
    rng = np.random.default_rng(21)
    a = _structured_rank_deficient(rng, m, n, r)
    q, rr = qb.qr(ctx.array(np.asfortranarray(a)), (0, 1), 1)
    q, rr = q.to_host(), rr.to_host()
    k = min(m, n)
    assert np.abs(q.conj().T @ q - np.eye(k)).max() < 1e-12          # orthonormal whatever the rank, as LAPACK's Q
    assert np.linalg.norm(q @ rr - a) <= 1e-12 * np.linalg.norm(a)
    assert np.abs(np.tril(rr[:, :k], -1)).max() < 1e-12 * np.abs(a).max()


@pytest.mark.parametrize("m,n,r", [(1280, 640, 200), (640, 1280, 256), (200, 120, 30)])
def test_svd_of_structured_rank_deficient_input(qb, ctx, m, n, r):
    rng = np.random.default_rng(22)
    a = _structured_rank_deficient(rng, max(m, n), min(m, n), r)
    if m < n:
        a = a.conj().T.copy()
    u, s, vc, kept, dw = qb.svd(ctx.array(np.asfortranarray(a)), (0, 1), 1)
    u, s, vc = u.to_host(), s.to_host(), vc.to_host()
    k = min(m, n)
    s_ref = sla.svd(a, compute_uv=False, lapack_driver="gesdd")
    assert kept == k and np.abs(s - s_ref).max() <= 1e-12 * s_ref[0]
    assert np.abs(u.conj().T @ u - np.eye(k)).max() < 1e-11          # null-space columns included
    assert np.abs(vc.T @ vc.conj() - np.eye(k)).max() < 1e-11
    # reconstruction: a few 1e-12 |A| on this pathological input (exact repetitions: cond = inf); 1e-15 for full rank
    assert np.linalg.norm((u * s) @ vc.T - a) <= 1e-11 * np.linalg.norm(a)
    # with the truncate! rule's absolute threshold the kept count is the rank
    _, s2, _, kept2, _ = qb.svd(ctx.array(np.asfortranarray(a)), (0, 1), 1, threshold=1e-10 * s_ref[0])
    assert kept2 == r


def test_canonize_of_an_mpo_applied_state_keeps_the_state(qb, ctx):
    """H|psi> has structurally rank-deficient site tensors (bond chi*D, rank < chi*D): canonize! must leave the state and
    its norm alone (it changed the norm by 23 % at n = 20, chi = 512 before the QR completed dependent columns)."""
    n, chi = 14, 128
    arrays = qb.rand_mps_arrays(np.random.default_rng(1003), n, chi)
    g = qb.B200MPS(ctx, arrays).apply_mpo(qb.heisenberg_mpo_arrays(n))
    n0 = g.norm()
    c = g.copy().canonize()
    assert abs(c.norm() - n0) <= 1e-10 * n0
    assert abs(c.overlap(g) - n0 ** 2) <= 1e-10 * n0 ** 2
    for lam in c.lambdas():
        assert abs(np.sum(lam ** 2) - n0 ** 2) <= 1e-10 * n0 ** 2
