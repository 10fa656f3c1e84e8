// Fused MPS path: canonize!, mixed_canonize!, truncate!, evolve!, overlap, expect of
// /root/reference/src/Ansatz/Chain.jl on a device-resident open-boundary MPS.
//
// Private layout: site tensor (l, o, r) column-major, so that
//   - the (l,o | r) matricisation used by left-canonisation / SVD sweeps is the array itself,
//   - the (l | o,r) matricisation used by right-canonisation is the array itself with ld = chi_l,
//   - theta = (Λl Γl Λ)(Γr Λr) is one plain GEMM with no permutation and comes out as
//     (l,o1) x (o2,r), which is what the gate kernel and the SVD consume.
// A Schmidt vector Λ_b (bond b between sites b and b+1) is a diagonal factor sitting ON the bond
// (Tenet hyper-index semantics): the state is  A_0 Λ_0 A_1 Λ_1 ... wherever Λ_b is present.
#include <algorithm>
#include <cmath>

#include "common.cuh"

#include <cstdlib>
#include <array>
#include <condition_variable>
#include <mutex>
#include <thread>

struct qb200_mps {
    int n = 0;
    int form = 0;  // 0 plain, 1 Vidal (after canonize!), 2 mixed
    // Storage type of the site tensors in HBM: QB200_C128, or QB200_C64 (float2: half the footprint).  The fused chains
    // compute in FP64 by explicit choice -- 95 % of the work is the Jacobi SVD, whose rotations need FP64 to deliver the
    // Schmidt values to 1e-12 -- so a ComplexF32 chain is widened for the duration of a call (WideScope) and its
    // results are rounded back to ComplexF32 once, at the end (parity bar: 1e-5, north star).
    int dtype = QB200_C128;
    bool wide = false;  // a C64 chain whose site pointers currently hold the widened (ComplexF64) copies
    std::vector<c128*> site;  // float2* in disguise while dtype == C64 && !wide
    std::vector<int64_t> chil, p, chir;
    std::vector<double*> lam;  // device, null when absent
    std::vector<std::vector<double>> lam_host;
};

namespace {

const c128 ONE = {1.0, 0.0}, ZERO = {0.0, 0.0};

__global__ void pinv_kernel(const double* __restrict__ lam, double* __restrict__ inv, int64_t n, double atol) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double v = lam[i];
        inv[i] = (fabs(v) > atol) ? 1.0 / v : 0.0;
    }
}

int32_t set_site_dev(qb200_ctx* ctx, qb200_mps* m, int s, c128* data, int64_t chil, int64_t p, int64_t chir) {
    if (m->site[s]) cudaFreeAsync(m->site[s], ctx->stream);
    m->site[s] = data;
    m->chil[s] = chil;
    m->p[s] = p;
    m->chir[s] = chir;
    return QB200_OK;
}

c128* dev_alloc(qb200_ctx* ctx, int64_t n) {
    void* p = nullptr;
    if (cudaMallocAsync(&p, sizeof(c128) * (size_t)std::max<int64_t>(n, 1), ctx->stream) != cudaSuccess) return nullptr;
    return (c128*)p;
}

int32_t set_lambda_host(qb200_ctx* ctx, qb200_mps* m, int b, const double* host, int64_t n) {
    if (m->lam[b]) cudaFreeAsync(m->lam[b], ctx->stream);
    m->lam[b] = nullptr;
    void* d = nullptr;
    QB_CUDA(ctx, cudaMallocAsync(&d, sizeof(double) * (size_t)std::max<int64_t>(n, 1), ctx->stream));
    m->lam[b] = (double*)d;
    m->lam_host[b].assign(host, host + n);
    QB_CUDA(ctx, cudaMemcpyAsync(d, m->lam_host[b].data(), sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    return QB200_OK;
}

void drop_lambda(qb200_ctx* ctx, qb200_mps* m, int b) {
    if (m->lam[b]) cudaFreeAsync(m->lam[b], ctx->stream);
    m->lam[b] = nullptr;
    m->lam_host[b].clear();
}

// truncate! rule (Chain.jl:404-417)
int64_t kept_count(const std::vector<double>& s, int64_t maxdim, double threshold) {
    int64_t k = (int64_t)s.size();
    int64_t lim = (maxdim > 0) ? std::min(k, maxdim) : k;
    int64_t kept = 0;
    for (int64_t i = 0; i < lim; ++i) {
        if (threshold < 0.0 || std::fabs(s[i]) > threshold)
            kept++;
        else
            break;
    }
    return kept;
}

// site s (chil*p x chir) = Q R, site s <- Q, site s+1 <- R * site s+1            [direction :right, :qr]
int32_t left_canonize_qr(qb200_ctx* ctx, qb200_mps* m, int s) {
    int64_t rows = m->chil[s] * m->p[s], cols = m->chir[s], k = std::min(rows, cols);
    c128* Q = dev_alloc(ctx, rows * k);
    Workspace ws(ctx);
    c128* R = ws.get<c128>((size_t)(k * cols));
    int64_t ncols = m->p[s + 1] * m->chir[s + 1];
    c128* nxt = dev_alloc(ctx, k * ncols);
    if (!Q || !R || !nxt) QB_FAIL(ctx, QB200_E_CUDA, "mps: out of device memory");
    QB_TRY(qb_qr_matrix(ctx, rows, cols, m->site[s], rows, Q, rows, R, k));
    QB_TRY(qb_gemm(ctx, 0, 0, k, ncols, cols, ONE, R, k, m->site[s + 1], cols, ZERO, nxt, k));
    set_site_dev(ctx, m, s, Q, m->chil[s], m->p[s], k);
    set_site_dev(ctx, m, s + 1, nxt, k, m->p[s + 1], m->chir[s + 1]);
    return QB200_OK;
}

// site s as chil x (p*chir): M = R^H Q^H with M^H = Q R; site s <- Q^H, site s-1 <- site s-1 * R^H   [:left, :qr]
int32_t right_canonize_qr(qb200_ctx* ctx, qb200_mps* m, int s) {
    int64_t rows = m->chil[s], cols = m->p[s] * m->chir[s], k = std::min(rows, cols);
    Workspace ws(ctx);
    c128* Mh = ws.get<c128>((size_t)(rows * cols));  // cols x rows
    c128* Q = ws.get<c128>((size_t)(cols * k));
    c128* R = ws.get<c128>((size_t)(k * rows));
    c128* Qh = dev_alloc(ctx, k * cols);
    int64_t prow = m->chil[s - 1] * m->p[s - 1];
    c128* prv = dev_alloc(ctx, prow * k);
    if (!Mh || !Q || !R || !Qh || !prv) QB_FAIL(ctx, QB200_E_CUDA, "mps: out of device memory");
    QB_TRY(qb_copy_matrix(ctx, rows, cols, m->site[s], rows, Mh, cols, 1));
    QB_TRY(qb_qr_matrix(ctx, cols, rows, Mh, cols, Q, cols, R, k));
    QB_TRY(qb_copy_matrix(ctx, cols, k, Q, cols, Qh, k, 1));
    QB_TRY(qb_gemm(ctx, 0, 2, prow, k, rows, ONE, m->site[s - 1], prow, R, k, ZERO, prv, prow));
    set_site_dev(ctx, m, s, Qh, k, m->p[s], m->chir[s]);
    set_site_dev(ctx, m, s - 1, prv, m->chil[s - 1], m->p[s - 1], k);
    return QB200_OK;
}

// site s (chil*p x chir) = U S V^H; site s <- U, Λ_s <- S, site s+1 <- V^H * site s+1  (S left on the bond)
int32_t left_canonize_svd(qb200_ctx* ctx, qb200_mps* m, int s, int64_t maxdim = 0, double threshold = -1.0) {
    int64_t rows = m->chil[s] * m->p[s], cols = m->chir[s], k = std::min(rows, cols);
    SvdState* st = nullptr;
    std::vector<double> sigma;
    QB_TRY(qb_svd_factor(ctx, rows, cols, m->site[s], rows, &st, sigma));
    if (maxdim > 0 || threshold >= 0.0) {  // truncate! right after the SVD (Chain.jl:404-417)
        k = kept_count(sigma, maxdim, threshold >= 0.0 ? threshold : 1e-16);
        if (k == 0) {
            qb_svd_release(ctx, st);
            QB_FAIL(ctx, QB200_E_INVALID, "compress: every Schmidt coefficient is below the threshold");
        }
        sigma.resize(k);
    }
    c128* U = dev_alloc(ctx, rows * k);
    Workspace ws(ctx);
    c128* Vh = ws.get<c128>((size_t)(k * cols));
    void* S = nullptr;
    cudaMallocAsync(&S, sizeof(double) * k, ctx->stream);
    int64_t ncols = m->p[s + 1] * m->chir[s + 1];
    c128* nxt = dev_alloc(ctx, k * ncols);
    if (!U || !Vh || !S || !nxt) {
        qb_svd_release(ctx, st);
        QB_FAIL(ctx, QB200_E_CUDA, "mps: out of device memory");
    }
    int32_t r = qb_svd_emit(ctx, st, k, U, rows, (double*)S, Vh, k, 1, nullptr, 0, nullptr, 1, 1.0);
    qb_svd_release(ctx, st);
    QB_TRY(r);
    QB_TRY(qb_gemm(ctx, 0, 0, k, ncols, cols, ONE, Vh, k, m->site[s + 1], cols, ZERO, nxt, k));
    set_site_dev(ctx, m, s, U, m->chil[s], m->p[s], k);
    set_site_dev(ctx, m, s + 1, nxt, k, m->p[s + 1], m->chir[s + 1]);
    drop_lambda(ctx, m, s);
    m->lam[s] = (double*)S;
    m->lam_host[s] = sigma;
    return QB200_OK;
}

// site s as chil x (p*chir) = U S V^H; site s <- V^H, Λ_{s-1} <- S, site s-1 <- site s-1 * U   [:left, :svd]
int32_t right_canonize_svd(qb200_ctx* ctx, qb200_mps* m, int s) {
    int64_t rows = m->chil[s], cols = m->p[s] * m->chir[s], k = std::min(rows, cols);
    SvdState* st = nullptr;
    std::vector<double> sigma;
    QB_TRY(qb_svd_factor(ctx, rows, cols, m->site[s], rows, &st, sigma));
    Workspace ws(ctx);
    c128* U = ws.get<c128>((size_t)(rows * k));
    c128* Vh = dev_alloc(ctx, k * cols);
    void* S = nullptr;
    cudaMallocAsync(&S, sizeof(double) * k, ctx->stream);
    int64_t prow = m->chil[s - 1] * m->p[s - 1];
    c128* prv = dev_alloc(ctx, prow * k);
    if (!U || !Vh || !S || !prv) {
        qb_svd_release(ctx, st);
        QB_FAIL(ctx, QB200_E_CUDA, "mps: out of device memory");
    }
    int32_t r = qb_svd_emit(ctx, st, k, U, rows, (double*)S, Vh, k, 1, nullptr, 0, nullptr, 1, 1.0);
    qb_svd_release(ctx, st);
    QB_TRY(r);
    QB_TRY(qb_gemm(ctx, 0, 0, prow, k, rows, ONE, m->site[s - 1], prow, U, rows, ZERO, prv, prow));
    set_site_dev(ctx, m, s, Vh, k, m->p[s], m->chir[s]);
    set_site_dev(ctx, m, s - 1, prv, m->chil[s - 1], m->p[s - 1], k);
    drop_lambda(ctx, m, s - 1);
    m->lam[s - 1] = (double*)S;
    m->lam_host[s - 1] = sigma;
    return QB200_OK;
}

// E'(ra x rb) = sum_{la,lb,o} E(la,lb) A(la,o,ra) conj(B(lb,o,rb)), then bond Schmidt vectors applied
int32_t transfer_left(qb200_ctx* ctx, const c128* E, int64_t la, int64_t lb, const c128* A, int64_t p, int64_t ra,
                      const c128* B, int64_t rb, const double* lamA, const double* lamB, c128* Eout, c128* tmp) {
    // tmp(lb x (o,ra)) = E^T A
    QB_TRY(qb_gemm(ctx, 1, 0, lb, p * ra, la, ONE, E, la, A, la, ZERO, tmp, lb));
    // Eout(ra x rb) = tmp((lb,o) x ra)^T conj(B((lb,o) x rb))
    QB_TRY(qb_gemm(ctx, 1, 3, ra, rb, lb * p, ONE, tmp, lb * p, B, lb * p, ZERO, Eout, ra));
    if (lamA || lamB) QB_TRY(qb_scale_rows_cols(ctx, Eout, Eout, ra, rb, lamA, ra, lamB, 1));
    return QB200_OK;
}

// Rout(la x lb) = sum_{o,ra,rb} A(la,o,ra) R(ra,rb) conj(B(lb,o,rb)), then the Schmidt vectors of the bond to
// the LEFT of the site applied
int32_t transfer_right(qb200_ctx* ctx, const c128* R, int64_t ra, int64_t rb, const c128* A, int64_t la, int64_t p,
                       const c128* B, int64_t lb, const double* lamA, const double* lamB, c128* Rout, c128* tmp) {
    // tmp((la,o) x rb) = A((la,o) x ra) R
    QB_TRY(qb_gemm(ctx, 0, 0, la * p, rb, ra, ONE, A, la * p, R, ra, ZERO, tmp, la * p));
    // Rout(la x lb) = tmp(la x (o,rb)) B(lb x (o,rb))^H
    QB_TRY(qb_gemm(ctx, 0, 2, la, lb, p * rb, ONE, tmp, la, B, lb, ZERO, Rout, la));
    if (lamA || lamB) QB_TRY(qb_scale_rows_cols(ctx, Rout, Rout, la, lb, lamA, la, lamB, 1));
    return QB200_OK;
}

// wrap raw device memory as a tensor view for qb200_contract
qb200_tensor view3(c128* p, int64_t a, int64_t b, int64_t c) {
    qb200_tensor t;
    t.dtype = QB200_C128;
    t.user_dtype = QB200_C128;
    t.rank = 3;
    t.ext[0] = a;
    t.ext[1] = b;
    t.ext[2] = c;
    t.data = p;
    t.bytes = 0;
    t.owned = false;
    return t;
}
qb200_tensor view4(c128* p, int64_t a, int64_t b, int64_t c, int64_t d) {
    qb200_tensor t = view3(p, a, b, c);
    t.rank = 4;
    t.ext[3] = d;
    return t;
}
qb200_tensor view5(c128* p, int64_t a, int64_t b, int64_t c, int64_t d, int64_t e) {
    qb200_tensor t = view4(p, a, b, c, d);
    t.rank = 5;
    t.ext[4] = e;
    return t;
}

// Λ_b -> site b+1 for every bond that holds a Schmidt vector: the chain becomes plain (same state)
int32_t absorb_lambdas(qb200_ctx* ctx, qb200_mps* m) {
    for (int b = 0; b < m->n - 1; ++b)
        if (m->lam[b]) {
            int64_t l = m->chil[b + 1], rest = m->p[b + 1] * m->chir[b + 1];
            QB_TRY(qb_scale_mode_raw(ctx, m->site[b + 1], m->site[b + 1], 1, l, rest, m->lam[b], 0, 0.0));
            drop_lambda(ctx, m, b);
        }
    return QB200_OK;
}

// real (Float64 / Float32) host data -> ComplexF64 on the device (`rand(...; eltype = Float64)` is the reference's
// default, Chain.jl:226-227: real chains are accepted at the boundary and computed in complex arithmetic)
template <typename T>
__global__ void widen_real_kernel(const T* __restrict__ src, c128* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = make_double2((double)src[i], 0.0);
}

inline int64_t site_elems(const qb200_mps* m, int s) { return m->chil[s] * m->p[s] * m->chir[s]; }
inline size_t site_esz(const qb200_mps* m) { return (m->dtype == QB200_C64 && !m->wide) ? sizeof(float2) : sizeof(c128); }

// For the duration of one public call a ComplexF32 chain holds ComplexF64 copies of its sites (every kernel of the fused
// path is FP64); on exit the (possibly new) sites are rounded back to float2 storage.  No-op for ComplexF64 chains and
// for nested scopes.  `writeback = false` for calls that only read the chain: the float2 originals are kept aside and
// put back, nothing is rounded.
struct WideScope {
    qb200_ctx* ctx;
    qb200_mps* m;
    bool active = false, writeback;
    std::vector<c128*> kept;  // float2 originals (read-only scopes)
    int32_t status = QB200_OK;
    WideScope(qb200_ctx* c, const qb200_mps* mc, bool wb) : ctx(c), m(const_cast<qb200_mps*>(mc)), writeback(wb) {
        if (!m || m->dtype != QB200_C64 || m->wide) return;
        active = true;
        kept.assign(m->n, nullptr);
        for (int s = 0; s < m->n; ++s) {
            if (!m->site[s]) continue;
            const int64_t cnt = site_elems(m, s);
            c128* w = dev_alloc(ctx, cnt);
            if (!w || qb_widen_c64(ctx, m->site[s], w, cnt) != QB200_OK) {
                status = QB200_E_CUDA;
                ctx->err = "mps: out of device memory while widening a ComplexF32 chain";
                if (w) cudaFreeAsync(w, ctx->stream);
                continue;  // the site stays narrow; the caller sees `status` and backs out
            }
            kept[s] = m->site[s];
            m->site[s] = w;
        }
        m->wide = (status == QB200_OK);
        if (status != QB200_OK) {
            restore_failed();
        } else if (writeback) {  // the float2 originals are dead from here on: the call will produce the new ones
            for (auto& p : kept)
                if (p) {
                    cudaFreeAsync(p, ctx->stream);
                    p = nullptr;
                }
        }
    }
    void restore_failed() {  // partial widening: put every site back to float2
        for (int s = 0; s < m->n; ++s)
            if (kept.size() > (size_t)s && kept[s]) {
                cudaFreeAsync(m->site[s], ctx->stream);
                m->site[s] = kept[s];
                kept[s] = nullptr;
            }
        active = false;
    }
    ~WideScope() {
        if (!active) return;
        for (int s = 0; s < m->n; ++s) {
            if (!m->site[s]) continue;
            if (!writeback) {
                cudaFreeAsync(m->site[s], ctx->stream);
                m->site[s] = kept[s];
                continue;
            }
            const int64_t cnt = site_elems(m, s);
            void* nar = nullptr;
            if (cudaMallocAsync(&nar, sizeof(float2) * (size_t)std::max<int64_t>(cnt, 1), ctx->stream) == cudaSuccess &&
                qb_narrow_c128(ctx, m->site[s], nar, cnt) == QB200_OK) {
                cudaFreeAsync(m->site[s], ctx->stream);
                m->site[s] = (c128*)nar;
            } else {  // cannot round back: keep the chain usable by switching its storage type
                m->dtype = QB200_C128;
            }
        }
        m->wide = false;
    }
};
#define QB_WIDE(ctx, m, writeback)      \
    WideScope _wide_##m(ctx, m, writeback); \
    if (_wide_##m.status != QB200_OK) return _wide_##m.status

int32_t check_site(qb200_ctx* ctx, const qb200_mps* m, int s) {
    if (!m || s < 0 || s >= m->n) QB_FAIL(ctx, QB200_E_INVALID, "mps: site %d out of range", s);
    if (!m->site[s]) QB_FAIL(ctx, QB200_E_INVALID, "mps: site %d not set", s);
    return QB200_OK;
}

int32_t check_complete(qb200_ctx* ctx, const qb200_mps* m) {
    if (!m) QB_FAIL(ctx, QB200_E_INVALID, "mps: null handle");
    for (int s = 0; s < m->n; ++s) {
        QB_TRY(check_site(ctx, m, s));
        if (s > 0 && m->chil[s] != m->chir[s - 1]) QB_FAIL(ctx, QB200_E_INVALID, "mps: bond %d dimension mismatch", s - 1);
    }
    if (m->chil[0] != 1 || m->chir[m->n - 1] != 1) QB_FAIL(ctx, QB200_E_INVALID, "mps: open boundary needs edge bonds of 1");
    return QB200_OK;
}

}  // namespace

extern "C" {

int32_t qb200_mps_create(qb200_ctx* ctx, int32_t nsites, qb200_mps** out) {
    if (!ctx || !out || nsites < 2) QB_FAIL(ctx, QB200_E_INVALID, "mps_create: need at least 2 sites");
    qb200_mps* m = new qb200_mps();
    m->n = nsites;
    m->site.assign(nsites, nullptr);
    m->chil.assign(nsites, 0);
    m->p.assign(nsites, 0);
    m->chir.assign(nsites, 0);
    m->lam.assign(nsites - 1, nullptr);
    m->lam_host.assign(nsites - 1, {});
    *out = m;
    return QB200_OK;
}

// storage type chosen at creation: QB200_C128 or QB200_C64 (ComplexF32 sites in HBM, FP64 arithmetic inside the calls)
int32_t qb200_mps_create_typed(qb200_ctx* ctx, int32_t nsites, int32_t dtype, qb200_mps** out) {
    if (dtype != QB200_C128 && dtype != QB200_C64)
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "mps_create: storage type must be ComplexF64 or ComplexF32 (real data is widened at set_site)");
    QB_TRY(qb200_mps_create(ctx, nsites, out));
    (*out)->dtype = dtype;
    return QB200_OK;
}
int32_t qb200_mps_dtype(const qb200_mps* m) { return m ? m->dtype : -1; }

int32_t qb200_mps_free(qb200_ctx* ctx, qb200_mps* m) {
    if (!m) return QB200_OK;
    for (auto p : m->site)
        if (p) cudaFreeAsync(p, ctx->stream);
    for (auto p : m->lam)
        if (p) cudaFreeAsync(p, ctx->stream);
    delete m;
    return QB200_OK;
}

int32_t qb200_mps_copy(qb200_ctx* ctx, const qb200_mps* src, qb200_mps** out) {
    if (!src || !out) QB_FAIL(ctx, QB200_E_INVALID, "mps_copy: null argument");
    qb200_mps* m = nullptr;
    QB_TRY(qb200_mps_create(ctx, src->n, &m));
    m->form = src->form;
    m->dtype = src->dtype;
    m->wide = src->wide;  // a copy taken inside a WideScope is a ComplexF64 scratch chain (never rounded back)
    const size_t esz = site_esz(src);
    for (int s = 0; s < src->n; ++s) {
        if (!src->site[s]) continue;
        int64_t cnt = src->chil[s] * src->p[s] * src->chir[s];
        c128* d = dev_alloc(ctx, (int64_t)((cnt * esz + sizeof(c128) - 1) / sizeof(c128)));
        if (!d) {
            qb200_mps_free(ctx, m);
            QB_FAIL(ctx, QB200_E_CUDA, "mps_copy: out of device memory");
        }
        cudaMemcpyAsync(d, src->site[s], esz * cnt, cudaMemcpyDeviceToDevice, ctx->stream);
        set_site_dev(ctx, m, s, d, src->chil[s], src->p[s], src->chir[s]);
    }
    for (int b = 0; b < src->n - 1; ++b) {
        if (!src->lam[b]) continue;
        size_t cnt = src->lam_host[b].size();
        void* d = nullptr;
        cudaMallocAsync(&d, sizeof(double) * std::max<size_t>(cnt, 1), ctx->stream);
        cudaMemcpyAsync(d, src->lam[b], sizeof(double) * cnt, cudaMemcpyDeviceToDevice, ctx->stream);
        m->lam[b] = (double*)d;
        m->lam_host[b] = src->lam_host[b];
    }
    *out = m;
    return QB200_OK;
}

// site (0-based) from a host array with extents (chi_l, p, chi_r) column-major of element type host_dtype
// (QB200_C128 / C64 / F64 / F32); converted on the device to the chain's storage type
int32_t qb200_mps_set_site_typed(qb200_ctx* ctx, qb200_mps* m, int32_t s, int32_t host_dtype, int64_t chil, int64_t p,
                                 int64_t chir, const void* host) {
    if (!m || s < 0 || s >= m->n || !host || chil < 1 || p < 1 || chir < 1)
        QB_FAIL(ctx, QB200_E_INVALID, "mps_set_site: bad argument");
    if (host_dtype < QB200_C128 || host_dtype > QB200_F32) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "mps_set_site: unknown element type");
    if (m->wide) QB_FAIL(ctx, QB200_E_INVALID, "mps_set_site: chain is in use");
    const int64_t cnt = chil * p * chir;
    const size_t hsz = dtype_size(host_dtype);
    Workspace ws(ctx);
    void* raw = ws.get<char>((size_t)cnt * hsz);
    if (!raw) QB_FAIL(ctx, QB200_E_CUDA, "mps_set_site: out of device memory");
    QB_CUDA(ctx, cudaMemcpyAsync(raw, host, hsz * cnt, cudaMemcpyHostToDevice, ctx->stream));
    const size_t ssz = (m->dtype == QB200_C64) ? sizeof(float2) : sizeof(c128);
    void* dst = nullptr;
    QB_CUDA(ctx, cudaMallocAsync(&dst, ssz * (size_t)cnt, ctx->stream));
    int32_t r = QB200_OK;
    if (host_dtype == m->dtype) {
        if (cudaMemcpyAsync(dst, raw, ssz * cnt, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) r = QB200_E_CUDA;
    } else {
        // through ComplexF64: real -> complex, ComplexF32 -> ComplexF64, then (for a ComplexF32 chain) one rounding
        c128* wide = (m->dtype == QB200_C128) ? (c128*)dst : ws.get<c128>((size_t)cnt);
        if (!wide) r = QB200_E_CUDA;
        const unsigned blocks = (unsigned)std::min<int64_t>((cnt + 255) / 256, 4096);
        if (r == QB200_OK) {
            if (host_dtype == QB200_C128) {
                if (wide != raw && cudaMemcpyAsync(wide, raw, sizeof(c128) * cnt, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess)
                    r = QB200_E_CUDA;
            } else if (host_dtype == QB200_C64) {
                r = qb_widen_c64(ctx, raw, wide, cnt);
            } else if (host_dtype == QB200_F64) {
                widen_real_kernel<double><<<blocks, 256, 0, ctx->stream>>>((const double*)raw, wide, cnt);
                ctx->launches++;
            } else {
                widen_real_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float*)raw, wide, cnt);
                ctx->launches++;
            }
        }
        if (r == QB200_OK && m->dtype == QB200_C64) r = qb_narrow_c128(ctx, wide, dst, cnt);
    }
    if (r == QB200_OK && qb_stream_sync(ctx) != cudaSuccess) r = QB200_E_CUDA;  // the host buffer is only borrowed
    if (r != QB200_OK) {
        cudaFreeAsync(dst, ctx->stream);
        if (ctx->err.empty()) ctx->err = "mps_set_site: conversion failed";
        return r;
    }
    return set_site_dev(ctx, m, s, (c128*)dst, chil, p, chir);
}

int32_t qb200_mps_set_site(qb200_ctx* ctx, qb200_mps* m, int32_t s, int64_t chil, int64_t p, int64_t chir,
                           const void* host) {
    return qb200_mps_set_site_typed(ctx, m, s, QB200_C128, chil, p, chir, host);
}

int32_t qb200_mps_site_dims(const qb200_mps* m, int32_t s, int64_t dims[3]) {
    if (!m || s < 0 || s >= m->n || !dims) return QB200_E_INVALID;
    dims[0] = m->chil[s];
    dims[1] = m->p[s];
    dims[2] = m->chir[s];
    return QB200_OK;
}

// site to the host as ComplexF64 (host_dtype = QB200_C128) or ComplexF32 (QB200_C64), whatever the storage type
int32_t qb200_mps_get_site_typed(qb200_ctx* ctx, const qb200_mps* m, int32_t s, int32_t host_dtype, void* host) {
    QB_TRY(check_site(ctx, m, s));
    if (!host || (host_dtype != QB200_C128 && host_dtype != QB200_C64)) QB_FAIL(ctx, QB200_E_INVALID, "mps_get_site: bad argument");
    const int64_t cnt = site_elems(m, s);
    const int have = (m->dtype == QB200_C64 && !m->wide) ? QB200_C64 : QB200_C128;
    Workspace ws(ctx);
    const void* src = m->site[s];
    if (have != host_dtype) {
        if (host_dtype == QB200_C128) {
            c128* w = ws.get<c128>((size_t)cnt);
            if (!w) QB_FAIL(ctx, QB200_E_CUDA, "mps_get_site: out of device memory");
            QB_TRY(qb_widen_c64(ctx, m->site[s], w, cnt));
            src = w;
        } else {
            float2* nar = ws.get<float2>((size_t)cnt);
            if (!nar) QB_FAIL(ctx, QB200_E_CUDA, "mps_get_site: out of device memory");
            QB_TRY(qb_narrow_c128(ctx, m->site[s], nar, cnt));
            src = nar;
        }
    }
    QB_CUDA(ctx, cudaMemcpyAsync(host, src, dtype_size(host_dtype) * cnt, cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    return QB200_OK;
}

int32_t qb200_mps_get_site(qb200_ctx* ctx, const qb200_mps* m, int32_t s, void* host) {
    return qb200_mps_get_site_typed(ctx, m, s, QB200_C128, host);
}

int32_t qb200_mps_set_lambda(qb200_ctx* ctx, qb200_mps* m, int32_t b, int64_t n, const double* host) {
    if (!m || b < 0 || b >= m->n - 1 || !host || n < 1) QB_FAIL(ctx, QB200_E_INVALID, "mps_set_lambda: bad argument");
    return set_lambda_host(ctx, m, b, host, n);
}

int32_t qb200_mps_get_lambda(qb200_ctx* ctx, const qb200_mps* m, int32_t b, double* host, int64_t* n) {
    if (!m || b < 0 || b >= m->n - 1) QB_FAIL(ctx, QB200_E_INVALID, "mps_get_lambda: bad bond");
    if (!m->lam[b]) {
        if (n) *n = 0;
        QB_FAIL(ctx, QB200_E_NOSPECTRUM, "Can't access the spectrum on bond (%d, %d)", b + 1, b + 2);
    }
    if (n) *n = (int64_t)m->lam_host[b].size();
    if (host) memcpy(host, m->lam_host[b].data(), sizeof(double) * m->lam_host[b].size());
    return QB200_OK;
}

int32_t qb200_mps_form(const qb200_mps* m) { return m ? m->form : -1; }
int32_t qb200_mps_set_form(qb200_mps* m, int32_t form) {
    if (!m || form < 0 || form > 2) return QB200_E_INVALID;
    m->form = form;
    return QB200_OK;
}

// canonize! (Chain.jl:469-497)
int32_t qb200_mps_canonize(qb200_ctx* ctx, qb200_mps* m) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, true);
    // Schmidt vectors already sitting on bonds are absorbed by the QR sweep's contract!(tn, virtualind) (Chain.jl:372)
    QB_TRY(absorb_lambdas(ctx, m));
    for (int s = m->n - 1; s >= 1; --s) QB_TRY(right_canonize_qr(ctx, m, s));
    for (int s = 0; s < m->n - 1; ++s) {
        QB_TRY(left_canonize_svd(ctx, m, s));
        // A_{s+1} <- Λ_s A_{s+1}   (Chain.jl:482-485)
        int64_t l = m->chil[s + 1], rest = m->p[s + 1] * m->chir[s + 1];
        QB_TRY(qb_scale_mode_raw(ctx, m->site[s + 1], m->site[s + 1], 1, l, rest, m->lam[s], 0, 0.0));
    }
    for (int s = 1; s < m->n; ++s) {
        // Γ_s = A_s Λ_{s-1}^{-1}, pinv atol 1e-64 (Chain.jl:488-494)
        int64_t l = m->chil[s], rest = m->p[s] * m->chir[s];
        QB_TRY(qb_scale_mode_raw(ctx, m->site[s], m->site[s], 1, l, rest, m->lam[s - 1], 1, 1e-64));
    }
    m->form = 1;
    return QB200_OK;
}

// mixed_canonize! (Chain.jl:509-524)
int32_t qb200_mps_mixed_canonize(qb200_ctx* ctx, qb200_mps* m, int32_t center) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, true);
    if (center < 1 || center >= m->n)
        QB_FAIL(ctx, QB200_E_INVALID, "Cannot right-canonize left-most tensor (center must be in 2..n)");
    // absorb any Schmidt vector into the site on its right: the chain becomes plain
    QB_TRY(absorb_lambdas(ctx, m));
    for (int s = 0; s < center; ++s) QB_TRY(left_canonize_qr(ctx, m, s));
    for (int s = m->n - 1; s > center; --s) QB_TRY(right_canonize_qr(ctx, m, s));
    QB_TRY(right_canonize_svd(ctx, m, center));
    m->form = 2;
    return QB200_OK;
}

// truncate! (Chain.jl:390-422)
int32_t qb200_mps_truncate(qb200_ctx* ctx, qb200_mps* m, int32_t b, int64_t maxdim, double threshold,
                           int64_t* kept_out) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, true);
    if (b < 0 || b >= m->n - 1) QB_FAIL(ctx, QB200_E_INVALID, "Invalid bond %d", b);
    if (!m->lam[b]) QB_FAIL(ctx, QB200_E_NOSPECTRUM, "Can't access the spectrum on bond (%d, %d)", b + 1, b + 2);
    if (maxdim <= 0 && threshold < 0.0) threshold = 1e-16;
    if (threshold < 0.0) threshold = 1e-16;  // reference default (Chain.jl:411-413)
    int64_t chi = (int64_t)m->lam_host[b].size();
    // the reference filters, it does not require a prefix; the spectrum is sorted so both agree
    int64_t kept = kept_count(m->lam_host[b], maxdim, threshold);
    if (kept_out) *kept_out = kept;
    if (kept == chi) return QB200_OK;
    if (kept == 0) QB_FAIL(ctx, QB200_E_INVALID, "truncate: every Schmidt coefficient is below the threshold");
    // left site: keep the first `kept` columns (contiguous prefix in (l,o,r) layout)
    int64_t lrows = m->chil[b] * m->p[b];
    c128* L = dev_alloc(ctx, lrows * kept);
    int64_t rcols = m->p[b + 1] * m->chir[b + 1];
    c128* R = dev_alloc(ctx, kept * rcols);
    if (!L || !R) QB_FAIL(ctx, QB200_E_CUDA, "truncate: out of device memory");
    QB_TRY(qb_copy_matrix(ctx, lrows, kept, m->site[b], lrows, L, lrows, 0));
    QB_TRY(qb_copy_matrix(ctx, kept, rcols, m->site[b + 1], chi, R, kept, 0));
    set_site_dev(ctx, m, b, L, m->chil[b], m->p[b], kept);
    set_site_dev(ctx, m, b + 1, R, kept, m->p[b + 1], m->chir[b + 1]);
    m->lam_host[b].resize(kept);  // device vector: prefix stays valid
    return QB200_OK;
}

int32_t qb200_mps_evolve1(qb200_ctx* ctx, qb200_mps* m, int32_t s, const void* gate) {
    QB_TRY(check_site(ctx, m, s));
    QB_WIDE(ctx, m, true);
    if (!gate) QB_FAIL(ctx, QB200_E_INVALID, "evolve1: null gate");
    int64_t p = m->p[s];
    if (p > QB200_MAX_PHYS) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "evolve1: physical dimension %lld > %d", (long long)p, QB200_MAX_PHYS);
    Workspace ws(ctx);
    c128* g = ws.get<c128>((size_t)(p * p));
    if (!g) QB_FAIL(ctx, QB200_E_CUDA, "evolve1: workspace allocation failed");
    memcpy(ctx->scratch_host, gate, sizeof(c128) * p * p);
    QB_CUDA(ctx, cudaMemcpyAsync(g, ctx->scratch_host, sizeof(c128) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    QB_TRY(qb_apply_gate1(ctx, m->site[s], m->chil[s], p, m->chir[s], g));
    QB_CUDA(ctx, qb_stream_sync(ctx));  // scratch_host is reused by later calls
    return QB200_OK;
}

}  // extern "C"

namespace {
// device buffers of one bond update that must not outlive a failed call
struct Evolve2Out {
    qb200_ctx* ctx;
    c128 *U = nullptr, *Vh = nullptr;
    void* S = nullptr;
    SvdState* st = nullptr;
    explicit Evolve2Out(qb200_ctx* c) : ctx(c) {}
    ~Evolve2Out() {
        if (st) qb_svd_release(ctx, st);
        if (U) cudaFreeAsync(U, ctx->stream);
        if (Vh) cudaFreeAsync(Vh, ctx->stream);
        if (S) cudaFreeAsync(S, ctx->stream);
    }
};
}  // namespace

// evolve!(ψ, gate; threshold, maxdim, iscanonical, renormalize) for a gate on sites (b, b+1)
// (evolve_2site!, contract_2sitewf!, unpack_2sitewf!: Chain.jl:606-722).
//   iscanonical != 0 (Chain.jl:615, :643): θ = Λl Γl Λ Γr Λr with the neighbouring Schmidt vectors where they exist
//       (contract_2sitewf!, :669-685), SVD, Γl = U Λl^-1, Γr = Λr^-1 V^H with pinv atol 1e-32 (unpack_2sitewf!, :693-722);
//       renormalize: Λ <- Λ / |Λ| (:653-654).
//   iscanonical == 0 (the reference's default): contract!(tn, bond) -- Γl, Γr and the Schmidt vector ON the bond if one
//       is there, nothing else (:615) --, svd! leaves U, s, V^H (:645); renormalize: normalize!(ψ, bond[1]) =
//       mixed_canonize!(ψ, sitel) + normalisation of the Schmidt vector left of sitel (:655-656, :532-536).
static int32_t evolve2_core(qb200_ctx* ctx, qb200_mps* m, int32_t b, const void* gate, int64_t maxdim,
                            double threshold, int32_t renormalize, int32_t iscanonical, int64_t* kept_out,
                            double* discarded_weight, bool validate) {
    if (validate) QB_TRY(check_complete(ctx, m));
    if (b < 0 || b >= m->n - 1) QB_FAIL(ctx, QB200_E_INVALID, "evolve2: bond %d out of range", b);
    if (!gate) QB_FAIL(ctx, QB200_E_INVALID, "evolve2: null gate");
    const int64_t p1 = m->p[b], p2 = m->p[b + 1], pp = p1 * p2;
    if (p1 > QB200_MAX_PHYS || p2 > QB200_MAX_PHYS || pp * pp * (int64_t)sizeof(c128) > 32768)
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "evolve2: physical dimensions %lld x %lld too large", (long long)p1, (long long)p2);
    const bool truncating = (maxdim > 0 || threshold >= 0.0);
    if (renormalize && truncating && !iscanonical && b == 0)  // mixed_canonize!(ψ, Site(1)) (Chain.jl:344)
        QB_FAIL(ctx, QB200_E_INVALID, "Cannot right-canonize left-most tensor");
    const bool vidal = iscanonical != 0;
    const int64_t chil = m->chil[b], chib = m->chir[b], chir = m->chir[b + 1];
    const int64_t rows = chil * p1, cols = p2 * chir;
    const double* laml = (vidal && b > 0) ? m->lam[b - 1] : nullptr;
    const double* lamr = (vidal && b + 1 < m->n - 1) ? m->lam[b + 1] : nullptr;
    const double* lamb = m->lam[b];  // the Schmidt vector on the bond itself is part of contract!(tn, bond) either way

    Workspace ws(ctx);
    c128* Al = ws.get<c128>((size_t)(rows * chib));
    c128* Br = ws.get<c128>((size_t)(chib * cols));
    c128* theta = ws.get<c128>((size_t)(rows * cols));
    c128* g = ws.get<c128>((size_t)(pp * pp));
    double* linv = ws.get<double>((size_t)chil);
    double* rinv = ws.get<double>((size_t)chir);
    if (!Al || !Br || !theta || !g || !linv || !rinv) QB_FAIL(ctx, QB200_E_CUDA, "evolve2: workspace allocation failed");
    memcpy(ctx->scratch_host, gate, sizeof(c128) * pp * pp);
    QB_CUDA(ctx, cudaMemcpyAsync(g, ctx->scratch_host, sizeof(c128) * pp * pp, cudaMemcpyHostToDevice, ctx->stream));

    // contract_2sitewf!: θ = (Λl Γl Λ)(Γr Λr)   (outer Λ's stay in the network, Chain.jl:679-682)
    const c128* Aop = m->site[b];
    const c128* Bop = m->site[b + 1];
    if (laml || lamb) {
        PhaseTimer pt(ctx, QB_PH_SCALE, 32.0 * rows * chib);
        QB_TRY(qb_scale_rows_cols(ctx, m->site[b], Al, rows, chib, laml, chil, lamb, 1));
        Aop = Al;
    }
    if (lamr) {
        PhaseTimer pt(ctx, QB_PH_SCALE, 32.0 * chib * cols);
        QB_TRY(qb_scale_rows_cols(ctx, m->site[b + 1], Br, chib, cols, nullptr, 1, lamr, p2));
        Bop = Br;
    }
    {
        PhaseTimer pt(ctx, QB_PH_THETA_GEMM, 8.0 * rows * cols * chib);
        QB_TRY(qb_gemm(ctx, 0, 0, rows, cols, chib, ONE, Aop, rows, Bop, chib, ZERO, theta, rows));
    }
    // gate on the two physical indices (Chain.jl:635-636)
    {
        PhaseTimer pt(ctx, QB_PH_GATE, 32.0 * rows * cols);
        if (p1 == 2 && p2 == 2) {
            QB_TRY(qb_apply_gate2(ctx, theta, chil, chir, g));
        } else {  // general physical dimensions: θ'[l,o1,o2,r] = Σ G[o1,o2,i1,i2] θ[l,i1,i2,r] as one K1 contraction
            c128* t2 = ws.get<c128>((size_t)(rows * cols));
            if (!t2) QB_FAIL(ctx, QB200_E_CUDA, "evolve2: workspace allocation failed");
            qb200_tensor T0 = view4(theta, chil, p1, p2, chir), G = view4(g, p1, p2, p1, p2), T1 = view4(t2, chil, p1, p2, chir);
            const int32_t mT0[4] = {0, 1, 2, 3}, mG[4] = {4, 5, 1, 2}, mT1[4] = {0, 4, 5, 3};
            QB_TRY(qb200_contract(ctx, &T0, mT0, 0, &G, mG, 0, &T1, mT1, nullptr, nullptr));
            theta = t2;
        }
    }
    // SVD (Chain.jl:705 / :645)
    Evolve2Out o(ctx);
    std::vector<double> sigma;
    {
        // algorithmic count of a thin complex SVD with U and V: 4 (14 m n^2 + 8 n^3), m >= n
        double mm = (double)std::max(rows, cols), nn = (double)std::min(rows, cols);
        PhaseTimer pt(ctx, QB_PH_SVD, 4.0 * (14.0 * mm * nn * nn + 8.0 * nn * nn * nn));
        QB_TRY(qb_svd_factor(ctx, rows, cols, theta, rows, &o.st, sigma));
    }
    int64_t k = (int64_t)sigma.size();
    int64_t kept = k;
    if (truncating) kept = kept_count(sigma, maxdim, threshold >= 0.0 ? threshold : 1e-16);
    if (kept == 0) QB_FAIL(ctx, QB200_E_INVALID, "evolve2: every Schmidt coefficient is below the threshold");
    double dw = 0.0, kw = 0.0;
    for (int64_t i = k - 1; i >= kept; --i) dw += sigma[i] * sigma[i];
    for (int64_t i = kept - 1; i >= 0; --i) kw += sigma[i] * sigma[i];
    double sscale = 1.0;
    if (renormalize && truncating && vidal && kw > 0.0) sscale = 1.0 / std::sqrt(kw);
    // unpack_2sitewf!: Γl = U Λl^-1, Γr = Λr^-1 V^H, pinv atol 1e-32 (Chain.jl:708-713)
    if (laml) {
        pinv_kernel<<<(unsigned)((chil + 255) / 256), 256, 0, ctx->stream>>>(laml, linv, chil, 1e-32);
        ctx->launches++;
    }
    if (lamr) {
        pinv_kernel<<<(unsigned)((chir + 255) / 256), 256, 0, ctx->stream>>>(lamr, rinv, chir, 1e-32);
        ctx->launches++;
    }
    o.U = dev_alloc(ctx, rows * kept);
    o.Vh = dev_alloc(ctx, kept * cols);
    cudaMallocAsync(&o.S, sizeof(double) * kept, ctx->stream);
    if (!o.U || !o.Vh || !o.S) QB_FAIL(ctx, QB200_E_CUDA, "evolve2: out of device memory");
    QB_TRY(qb_svd_emit(ctx, o.st, kept, o.U, rows, (double*)o.S, o.Vh, kept, 1, laml ? linv : nullptr, chil,
                       lamr ? rinv : nullptr, p2, sscale));
    QB_CUDA(ctx, qb_stream_sync(ctx));  // scratch_host (gate) is free again
    set_site_dev(ctx, m, b, o.U, chil, p1, kept);
    set_site_dev(ctx, m, b + 1, o.Vh, kept, p2, chir);
    drop_lambda(ctx, m, b);
    m->lam[b] = (double*)o.S;
    o.U = o.Vh = nullptr;  // ownership moved into the chain
    o.S = nullptr;
    sigma.resize(kept);
    for (auto& v : sigma) v *= sscale;
    m->lam_host[b] = sigma;
    if (!vidal) m->form = 0;  // U s V^H on the bond: no longer a Vidal chain
    if (renormalize && truncating && !vidal) {
        // normalize!(ψ, bond[1]) (Chain.jl:532-536): mixed-canonical form centred on sitel, its Schmidt vector normalised
        QB_TRY(qb200_mps_mixed_canonize(ctx, m, b));
        std::vector<double> lam = m->lam_host[b - 1];
        double n2 = 0.0;
        for (double v : lam) n2 += v * v;
        if (n2 > 0.0) {
            const double f = 1.0 / std::sqrt(n2);
            for (double& v : lam) v *= f;
            QB_TRY(set_lambda_host(ctx, m, b - 1, lam.data(), (int64_t)lam.size()));
        }
    }
    if (kept_out) *kept_out = kept;
    if (discarded_weight) *discarded_weight = dw;
    return QB200_OK;
}

extern "C" {

int32_t qb200_mps_evolve2(qb200_ctx* ctx, qb200_mps* m, int32_t b, const void* gate, int64_t maxdim, double threshold,
                          int32_t renormalize, int32_t iscanonical, int64_t* kept_out, double* discarded_weight) {
    if (!ctx || !m) QB_FAIL(ctx, QB200_E_INVALID, "evolve2: null argument");
    QB_WIDE(ctx, m, true);
    return evolve2_core(ctx, m, b, gate, maxdim, threshold, renormalize, iscanonical, kept_out, discarded_weight, true);
}

namespace {
// Dependency-scheduled pool of two-site updates.  ops are given in program order; op i must wait for every earlier
// op that touches one of its two sites (bonds b-1, b, b+1) -- updates further apart only READ the Schmidt vector
// that sits between them, so they commute and run concurrently on worker streams, which hides the latency-bound
// phases of one update (QR panels, the shared-memory Jacobi solves) behind the DMMA-bound phases of the others.
// Every op sees exactly the inputs it would see in a sequential run => results are bit-identical to it.
int32_t run_evolve2_ops(qb200_ctx* ctx, qb200_mps* m, int32_t nops, const int32_t* bonds, const c128* g, int64_t maxdim,
                        double threshold, int32_t renormalize, int32_t iscanonical, int64_t* kept_out,
                        double* discarded_weight) {
    // gate i holds (p_b p_{b+1})^2 numbers (16 for qubits); physical dimensions never change, so the offsets are static
    std::vector<size_t> goff(nops + 1, 0);
    for (int i = 0; i < nops; ++i) {
        const size_t pp = (size_t)(m->p[bonds[i]] * m->p[bonds[i] + 1]);
        goff[i + 1] = goff[i] + pp * pp;
    }
    // worker streams: 12 by default, fewer when the host is small for the number of ranks sharing it (the workers
    // sleep on blocking events, so a 2x oversubscription of the cores is harmless)
    int nworkers = 12;
    {
        int hw = (int)std::thread::hardware_concurrency();
        int lws = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) lws = std::max(1, atoi(e));
        if (hw > 0) nworkers = std::min(12, std::max(4, 2 * hw / lws));
    }
    if (const char* e = getenv("QB200_WORKERS")) nworkers = std::max(1, atoi(e));
    nworkers = std::min(nworkers, (int)nops);
    // normalize!(ψ, bond[1]) of the non-canonical branch re-canonizes the whole chain: such updates do not commute
    if (!iscanonical && renormalize && (maxdim > 0 || threshold >= 0.0)) nworkers = 1;
    std::vector<int64_t> kept_tmp(nops, 0);
    std::vector<double> dw_tmp(nops, 0.0);
    if (nworkers <= 1) {
        for (int i = 0; i < nops; ++i)
            QB_TRY(evolve2_core(ctx, m, bonds[i], g + goff[i], maxdim, threshold, renormalize, iscanonical, &kept_tmp[i],
                                &dw_tmp[i], false));
    } else {
        // dependencies: the latest earlier op on each of the bonds b-1, b, b+1
        std::vector<std::array<int, 3>> dep(nops);
        {
            std::vector<int> last(m->n + 1, -1);
            for (int i = 0; i < nops; ++i) {
                for (int d = 0; d < 3; ++d) {
                    int bb = bonds[i] - 1 + d;
                    dep[i][d] = (bb >= 0 && bb < m->n - 1) ? last[bb] : -1;
                }
                last[bonds[i]] = i;
            }
        }
        // workers start after everything already queued on the parent stream
        QB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        std::vector<qb200_ctx*> w(nworkers);
        for (int t = 0; t < nworkers; ++t) {
            w[t] = qb_worker(ctx, t);
            w[t]->prof_on = ctx->prof_on;
            QB_CUDA(ctx, cudaStreamWaitEvent(w[t]->stream, ctx->ev1, 0));
        }
        std::mutex mu;
        std::condition_variable cv;
        std::vector<char> state(nops, 0);  // 0 pending, 1 running, 2 done
        int remaining = nops;
        int32_t first_rc = QB200_OK;
        int failed_worker = -1;
        // among the ready ops: the largest theta first (longest job), ties in program order.  The sizes are a
        // snapshot taken before the workers start (they rewrite chil / chir of their own sites as they finish).
        std::vector<int64_t> cost_of(nops);
        for (int i = 0; i < nops; ++i) cost_of[i] = m->chil[bonds[i]] * m->chir[bonds[i] + 1];
        auto cost = [&](int i) { return cost_of[i]; };
        auto pick = [&]() {
            int best = -1;
            int64_t bc = -1;
            for (int i = 0; i < nops; ++i) {
                if (state[i] != 0) continue;
                bool ready = true;
                for (int d = 0; d < 3; ++d)
                    if (dep[i][d] >= 0 && state[dep[i][d]] != 2) ready = false;
                if (!ready) continue;
                int64_t c = cost(i);
                if (c > bc) { bc = c; best = i; }
            }
            return best;
        };
        std::vector<std::thread> threads;
        for (int t = 0; t < nworkers; ++t)
            threads.emplace_back([&, t]() {
                cudaSetDevice(ctx->device);
                std::unique_lock<std::mutex> lk(mu);
                for (;;) {
                    int i = -1;
                    cv.wait(lk, [&]() { return remaining == 0 || first_rc != QB200_OK || (i = pick()) >= 0; });
                    if (i < 0) break;
                    state[i] = 1;
                    lk.unlock();
                    // evolve2_core returns with its stream drained: the op is complete on the device
                    int32_t r = evolve2_core(w[t], m, bonds[i], g + goff[i], maxdim, threshold, renormalize, iscanonical,
                                             &kept_tmp[i], &dw_tmp[i], false);
                    lk.lock();
                    state[i] = 2;
                    --remaining;
                    if (r != QB200_OK && first_rc == QB200_OK) { first_rc = r; failed_worker = t; }
                    cv.notify_all();
                }
                lk.unlock();
                qb_stream_sync(w[t]);
            });
        for (auto& th : threads) th.join();
        if (first_rc != QB200_OK) {
            if (failed_worker >= 0 && !w[failed_worker]->err.empty()) ctx->err = w[failed_worker]->err;
            return first_rc;
        }
    }
    for (int i = 0; i < nops; ++i) {
        if (kept_out) kept_out[i] = kept_tmp[i];
        if (discarded_weight) discarded_weight[i] = dw_tmp[i];
    }
    return QB200_OK;
}
}  // namespace

// One TEBD layer: `nb` two-site gates on pairwise non-adjacent bonds (e.g. all odd or all even bonds): the updates
// commute (disjoint sites; the Schmidt vectors between them are only read) and run as concurrent independent units.
// Results are identical to calling evolve! bond by bond.
int32_t qb200_mps_evolve2_layer(qb200_ctx* ctx, qb200_mps* m, int32_t nb, const int32_t* bonds, const void* gates,
                                int64_t maxdim, double threshold, int32_t renormalize, int32_t iscanonical,
                                int64_t* kept_out, double* discarded_weight) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, true);
    if (nb < 0 || (nb > 0 && (!bonds || !gates))) QB_FAIL(ctx, QB200_E_INVALID, "evolve2_layer: bad argument");
    std::vector<int> sorted(bonds, bonds + nb);
    std::sort(sorted.begin(), sorted.end());
    for (int i = 0; i < nb; ++i) {
        if (sorted[i] < 0 || sorted[i] >= m->n - 1) QB_FAIL(ctx, QB200_E_INVALID, "evolve2_layer: bond out of range");
        if (i > 0 && sorted[i] - sorted[i - 1] < 2)
            QB_FAIL(ctx, QB200_E_INVALID, "evolve2_layer: bonds %d and %d overlap", sorted[i - 1], sorted[i]);
    }
    if (nb == 0) return QB200_OK;
    return run_evolve2_ops(ctx, m, nb, bonds, (const c128*)gates, maxdim, threshold, renormalize, iscanonical, kept_out,
                           discarded_weight);
}

// A gate list in program order (a circuit of nearest-neighbour two-site gates, e.g. several TEBD layers or whole
// sweeps): the user loop `for (G, bond) in circuit evolve!(psi, G; ...)`.  Bonds may repeat and touch; an update
// starts as soon as the earlier updates on its two sites are complete, so consecutive layers overlap and no worker
// idles at a layer boundary.  Results are identical to nops calls of qb200_mps_evolve2 in the given order.
int32_t qb200_mps_evolve2_circuit(qb200_ctx* ctx, qb200_mps* m, int32_t nops, const int32_t* bonds, const void* gates,
                                  int64_t maxdim, double threshold, int32_t renormalize, int32_t iscanonical,
                                  int64_t* kept_out, double* discarded_weight) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, true);
    if (nops < 0 || (nops > 0 && (!bonds || !gates))) QB_FAIL(ctx, QB200_E_INVALID, "evolve2_circuit: bad argument");
    for (int i = 0; i < nops; ++i)
        if (bonds[i] < 0 || bonds[i] >= m->n - 1) QB_FAIL(ctx, QB200_E_INVALID, "evolve2_circuit: bond out of range");
    if (nops == 0) return QB200_OK;
    return run_evolve2_ops(ctx, m, nops, bonds, (const c128*)gates, maxdim, threshold, renormalize, iscanonical,
                           kept_out, discarded_weight);
}

// canonize! with truncate! applied to each bond right after its SVD: the composition a user of the reference
// writes to compress a state (e.g. after an MPO application).  maxdim <= 0 and threshold < 0:
// plain canonize!.
int32_t qb200_mps_compress(qb200_ctx* ctx, qb200_mps* m, int64_t maxdim, double threshold) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, true);
    QB_TRY(absorb_lambdas(ctx, m));  // the sweeps below start from a plain chain
    for (int s = m->n - 1; s >= 1; --s) QB_TRY(right_canonize_qr(ctx, m, s));
    for (int s = 0; s < m->n - 1; ++s) {
        QB_TRY(left_canonize_svd(ctx, m, s, maxdim, threshold));
        int64_t l = m->chil[s + 1], rest = m->p[s + 1] * m->chir[s + 1];
        QB_TRY(qb_scale_mode_raw(ctx, m->site[s + 1], m->site[s + 1], 1, l, rest, m->lam[s], 0, 0.0));
    }
    for (int s = 1; s < m->n; ++s) {
        int64_t l = m->chil[s], rest = m->p[s] * m->chir[s];
        QB_TRY(qb_scale_mode_raw(ctx, m->site[s], m->site[s], 1, l, rest, m->lam[s - 1], 1, 1e-64));
    }
    m->form = 1;
    return QB200_OK;
}

namespace {
// upload the MPO sites (host, (o, i, l, r) column-major each, concatenated) to one device buffer
int32_t upload_mpo(qb200_ctx* ctx, const qb200_mps* m, const int64_t* dl, const int64_t* dr, const void* sites,
                   Workspace& ws, std::vector<c128*>* dev) {
    int64_t total = 0;
    for (int s = 0; s < m->n; ++s) {
        if (dl[s] < 1 || dr[s] < 1) QB_FAIL(ctx, QB200_E_INVALID, "mpo: bad bond dimension at site %d", s);
        if (s > 0 && dl[s] != dr[s - 1]) QB_FAIL(ctx, QB200_E_INVALID, "mpo: bond %d dimension mismatch", s - 1);
        total += m->p[s] * m->p[s] * dl[s] * dr[s];
    }
    if (dl[0] != 1 || dr[m->n - 1] != 1) QB_FAIL(ctx, QB200_E_INVALID, "mpo: open boundary needs edge bonds of 1");
    c128* buf = ws.get<c128>((size_t)total);
    if (!buf) QB_FAIL(ctx, QB200_E_CUDA, "mpo: workspace allocation failed");
    QB_CUDA(ctx, cudaMemcpyAsync(buf, sites, sizeof(c128) * total, cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    int64_t off = 0;
    dev->clear();
    for (int s = 0; s < m->n; ++s) {
        dev->push_back(buf + off);
        off += m->p[s] * m->p[s] * dl[s] * dr[s];
    }
    return QB200_OK;
}
}  // namespace

// MPO application, site by site: B_s[(la,lw), o, (ra,rw)] = sum_i W_s[o,i,lw,rw] A_s[la,i,ra]
// (= contract(merge(Quantum(ψ), Quantum(H))) over the physical indices; bonds fuse with the ψ bond fastest).
// Schmidt vectors are absorbed first; the result is a plain chain with bonds chi*D -- call qb200_mps_compress.
int32_t qb200_mps_apply_mpo(qb200_ctx* ctx, qb200_mps* m, const int64_t* dl, const int64_t* dr, const void* sites) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, true);
    if (!dl || !dr || !sites) QB_FAIL(ctx, QB200_E_INVALID, "apply_mpo: null argument");
    Workspace ws(ctx);
    std::vector<c128*> W;
    QB_TRY(upload_mpo(ctx, m, dl, dr, sites, ws, &W));
    QB_TRY(absorb_lambdas(ctx, m));
    const int32_t mA[3] = {0, 1, 2};        // la, i, ra
    const int32_t mW[4] = {3, 1, 4, 5};     // o, i, lw, rw
    const int32_t mC[5] = {0, 4, 3, 2, 5};  // la, lw, o, ra, rw
    for (int s = 0; s < m->n; ++s) {
        int64_t la = m->chil[s], p = m->p[s], ra = m->chir[s];
        c128* out = dev_alloc(ctx, la * dl[s] * p * ra * dr[s]);
        if (!out) QB_FAIL(ctx, QB200_E_CUDA, "apply_mpo: out of device memory");
        qb200_tensor A = view3(m->site[s], la, p, ra), Wt = view4(W[s], p, p, dl[s], dr[s]);
        qb200_tensor C = view5(out, la, dl[s], p, ra, dr[s]);
        int32_t r = qb200_contract(ctx, &A, mA, 0, &Wt, mW, 0, &C, mC, nullptr, nullptr);
        if (r != QB200_OK) {
            cudaFreeAsync(out, ctx->stream);
            return r;
        }
        set_site_dev(ctx, m, s, out, la * dl[s], p, ra * dr[s]);
    }
    m->form = 0;
    return QB200_OK;
}

// <ψ|H|ψ> = contract(merge(ψ, H, ψ')), un-normalised, by a left-environment sweep L[a, w, b] resident in HBM
int32_t qb200_mps_expect_mpo(qb200_ctx* ctx, const qb200_mps* m, const int64_t* dl, const int64_t* dr,
                             const void* sites, double result[2]) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, false);
    if (!dl || !dr || !sites || !result) QB_FAIL(ctx, QB200_E_INVALID, "expect_mpo: null argument");
    Workspace ws(ctx);
    std::vector<c128*> W;
    QB_TRY(upload_mpo(ctx, m, dl, dr, sites, ws, &W));
    int64_t chimax = 1, dmax = 1, pmax = 1;
    for (int s = 0; s < m->n; ++s) {
        chimax = std::max(chimax, std::max(m->chil[s], m->chir[s]));
        dmax = std::max(dmax, std::max(dl[s], dr[s]));
        pmax = std::max(pmax, m->p[s]);
    }
    c128* L0 = ws.get<c128>((size_t)(chimax * dmax * chimax));
    c128* L1 = ws.get<c128>((size_t)(chimax * dmax * chimax));
    c128* T1 = ws.get<c128>((size_t)(dmax * chimax * pmax * chimax));
    c128* T2 = ws.get<c128>((size_t)(chimax * chimax * pmax * dmax));
    c128* As = ws.get<c128>((size_t)(chimax * pmax * chimax));
    if (!L0 || !L1 || !T1 || !T2 || !As) QB_FAIL(ctx, QB200_E_CUDA, "expect_mpo: workspace allocation failed");
    ctx->scratch_host[0] = 1.0;
    ctx->scratch_host[1] = 0.0;
    QB_CUDA(ctx, cudaMemcpyAsync(L0, ctx->scratch_host, sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    // modes: a=0 (ket left), w=1 (mpo left), b=2 (bra left), i=3, ra=4, o=5, rw=6, rb=7
    const int32_t mL[3] = {0, 1, 2}, mA[3] = {0, 3, 4}, mT1[4] = {1, 2, 3, 4};
    const int32_t mW[4] = {5, 3, 1, 6}, mT2[4] = {2, 4, 5, 6};
    const int32_t mB[3] = {2, 5, 7}, mLn[3] = {4, 6, 7};
    for (int s = 0; s < m->n; ++s) {
        int64_t la = m->chil[s], p = m->p[s], ra = m->chir[s];
        // effective site: the Schmidt vector of the bond to the right (if any) absorbed
        const c128* site = m->site[s];
        if (s < m->n - 1 && m->lam[s]) {
            QB_TRY(qb_scale_mode_raw(ctx, m->site[s], As, la * p, ra, 1, m->lam[s], 0, 0.0));
            site = As;
        }
        qb200_tensor L = view3(L0, la, dl[s], la), A = view3((c128*)site, la, p, ra);
        qb200_tensor t1 = view4(T1, dl[s], la, p, ra);
        QB_TRY(qb200_contract(ctx, &L, mL, 0, &A, mA, 0, &t1, mT1, nullptr, nullptr));
        qb200_tensor Wt = view4(W[s], p, p, dl[s], dr[s]), t2 = view4(T2, la, ra, p, dr[s]);
        QB_TRY(qb200_contract(ctx, &t1, mT1, 0, &Wt, mW, 0, &t2, mT2, nullptr, nullptr));
        qb200_tensor Ln = view3(L1, ra, dr[s], ra);
        QB_TRY(qb200_contract(ctx, &t2, mT2, 0, &A, mB, 1, &Ln, mLn, nullptr, nullptr));
        std::swap(L0, L1);
    }
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, L0, sizeof(c128), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    result[0] = ctx->scratch_host[0];
    result[1] = ctx->scratch_host[1];
    return QB200_OK;
}

// overlap(a, b) = <b|a> (Chain.jl:737-748): left-environment sweep
int32_t qb200_mps_overlap(qb200_ctx* ctx, const qb200_mps* a, const qb200_mps* b, double result[2]) {
    QB_TRY(check_complete(ctx, a));
    QB_TRY(check_complete(ctx, b));
    QB_WIDE(ctx, a, false);
    QB_WIDE(ctx, b, false);
    if (a->n != b->n) QB_FAIL(ctx, QB200_E_INVALID, "Ansatzes must have the same sites");
    int64_t maxa = 1, maxb = 1, maxp = 1;
    for (int s = 0; s < a->n; ++s) {
        if (a->p[s] != b->p[s]) QB_FAIL(ctx, QB200_E_INVALID, "overlap: physical dimensions differ at site %d", s);
        maxa = std::max(maxa, std::max(a->chil[s], a->chir[s]));
        maxb = std::max(maxb, std::max(b->chil[s], b->chir[s]));
        maxp = std::max(maxp, a->p[s]);
    }
    Workspace ws(ctx);
    c128* E0 = ws.get<c128>((size_t)(maxa * maxb));
    c128* E1 = ws.get<c128>((size_t)(maxa * maxb));
    c128* tmp = ws.get<c128>((size_t)(maxb * maxp * maxa));
    if (!E0 || !E1 || !tmp) QB_FAIL(ctx, QB200_E_CUDA, "overlap: workspace allocation failed");
    ctx->scratch_host[0] = 1.0;
    ctx->scratch_host[1] = 0.0;
    QB_CUDA(ctx, cudaMemcpyAsync(E0, ctx->scratch_host, sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    for (int s = 0; s < a->n; ++s) {
        const double* la = (s < a->n - 1) ? a->lam[s] : nullptr;
        const double* lb = (s < b->n - 1) ? b->lam[s] : nullptr;
        QB_TRY(transfer_left(ctx, E0, a->chil[s], b->chil[s], a->site[s], a->p[s], a->chir[s], b->site[s], b->chir[s],
                             la, lb, E1, tmp));
        std::swap(E0, E1);
    }
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, E0, sizeof(c128), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    result[0] = ctx->scratch_host[0];
    result[1] = ctx->scratch_host[1];
    return QB200_OK;
}

// expect(ψ, [O_s]) for a batch of single-site observables (Chain.jl:724-735, un-normalised):
// all left environments L_s and right environments R_s stay resident in HBM and are shared by the batch.
int32_t qb200_mps_expect1_batch(qb200_ctx* ctx, const qb200_mps* m, int32_t nobs, const int32_t* sites,
                                const void* ops, double* results) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, false);
    if (nobs < 0 || (nobs > 0 && (!sites || !ops || !results))) QB_FAIL(ctx, QB200_E_INVALID, "expect: bad argument");
    if (nobs == 0) return QB200_OK;
    const int n = m->n;
    int smin = n, smax = -1;
    int64_t pmax = 1, chimax = 1;
    for (int i = 0; i < nobs; ++i) {
        if (sites[i] < 0 || sites[i] >= n) QB_FAIL(ctx, QB200_E_INVALID, "expect: site %d out of range", sites[i]);
        smin = std::min(smin, (int)sites[i]);
        smax = std::max(smax, (int)sites[i]);
    }
    for (int s = 0; s < n; ++s) {
        pmax = std::max(pmax, m->p[s]);
        chimax = std::max(chimax, std::max(m->chil[s], m->chir[s]));
    }
    Workspace ws(ctx);
    // L[s]: environment to the left of site s (chil x chil), for s in [0, smax]; R[s]: to the right of site s
    // (chir x chir), for s in [smin, n-1]
    std::vector<c128*> L(n, nullptr), R(n, nullptr);
    c128* tmp = ws.get<c128>((size_t)(chimax * pmax * chimax));
    c128* tmp2 = ws.get<c128>((size_t)(chimax * pmax * chimax));
    c128* gates = ws.get<c128>((size_t)(nobs * pmax * pmax));
    c128* res = ws.get<c128>((size_t)nobs);
    if (!tmp || !tmp2 || !gates || !res) QB_FAIL(ctx, QB200_E_CUDA, "expect: workspace allocation failed");
    for (int s = 0; s <= smax; ++s) {
        L[s] = ws.get<c128>((size_t)(m->chil[s] * m->chil[s]));
        if (!L[s]) QB_FAIL(ctx, QB200_E_CUDA, "expect: workspace allocation failed");
    }
    for (int s = smin; s < n; ++s) {
        R[s] = ws.get<c128>((size_t)(m->chir[s] * m->chir[s]));
        if (!R[s]) QB_FAIL(ctx, QB200_E_CUDA, "expect: workspace allocation failed");
    }
    ctx->scratch_host[0] = 1.0;
    ctx->scratch_host[1] = 0.0;
    QB_CUDA(ctx, cudaMemcpyAsync(L[0], ctx->scratch_host, sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, cudaMemcpyAsync(R[n - 1], ctx->scratch_host, sizeof(c128), cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    for (int s = 0; s < smax; ++s) {
        const double* l = m->lam[s];
        QB_TRY(transfer_left(ctx, L[s], m->chil[s], m->chil[s], m->site[s], m->p[s], m->chir[s], m->site[s], m->chir[s],
                             l, l, L[s + 1], tmp));
    }
    for (int s = n - 1; s > smin; --s) {
        // R[s-1] = transfer over site s of R[s], with Λ_{s-1} (the bond left of site s) applied
        const double* l = m->lam[s - 1];
        QB_TRY(transfer_right(ctx, R[s], m->chir[s], m->chir[s], m->site[s], m->chil[s], m->p[s], m->site[s],
                              m->chil[s], l, l, R[s - 1], tmp));
    }
    // upload the operators in one go (pinned staging in chunks of the scratch buffer)
    {
        const c128* src = (const c128*)ops;
        // ops are given as nobs blocks of p_s * p_s numbers, concatenated
        int64_t off = 0;
        for (int i = 0; i < nobs; ++i) {
            int64_t p = m->p[sites[i]];
            if (p > QB200_MAX_PHYS)
                QB_FAIL(ctx, QB200_E_UNSUPPORTED, "expect: physical dimension %lld > %d", (long long)p, QB200_MAX_PHYS);
            memcpy(ctx->scratch_host, src + off, sizeof(c128) * p * p);
            QB_CUDA(ctx, cudaMemcpyAsync(gates + (size_t)i * pmax * pmax, ctx->scratch_host, sizeof(c128) * p * p,
                                         cudaMemcpyHostToDevice, ctx->stream));
            QB_CUDA(ctx, qb_stream_sync(ctx));
            off += p * p;
        }
    }
    for (int i = 0; i < nobs; ++i) {
        int s = sites[i];
        int64_t la = m->chil[s], p = m->p[s], ra = m->chir[s];
        // ϕ_s = O A_s (evolve_1site!), E = transfer_left(L_s, ϕ_s, A_s) without Λ, result = sum E .* R_s
        QB_CUDA(ctx, cudaMemcpyAsync(tmp2, m->site[s], sizeof(c128) * la * p * ra, cudaMemcpyDeviceToDevice, ctx->stream));
        QB_TRY(qb_apply_gate1(ctx, tmp2, la, p, ra, gates + (size_t)i * pmax * pmax));
        Workspace ws2(ctx);
        c128* Eo = ws2.get<c128>((size_t)(ra * ra));
        if (!Eo) QB_FAIL(ctx, QB200_E_CUDA, "expect: workspace allocation failed");
        QB_TRY(transfer_left(ctx, L[s], la, la, tmp2, p, ra, m->site[s], ra, nullptr, nullptr, Eo, tmp));
        // <E, R> = sum_{ra,rb} E(ra,rb) R(ra,rb): 1 x 1 GEMM with K = ra*ra (split-K reduction)
        QB_TRY(qb_gemm(ctx, 1, 0, 1, 1, ra * ra, ONE, Eo, ra * ra, R[s], ra * ra, ZERO, res + i, 1));
    }
    std::vector<c128> host(nobs);
    QB_CUDA(ctx, cudaMemcpyAsync(host.data(), res, sizeof(c128) * nobs, cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    for (int i = 0; i < nobs; ++i) {
        results[2 * i] = host[i].x;
        results[2 * i + 1] = host[i].y;
    }
    return QB200_OK;
}

// Replicate a device-resident MPS on every rank of the communicator (batched independent expectation values,
// "MPS replicated (one ncclBroadcast, <= 1.45 GiB); observables dealt round-robin").  On `root`
// *inout is the chain to send; on the other ranks *inout must be NULL and receives a new handle.  Shapes travel
// first (one small broadcast), then every site tensor and Schmidt vector in place over NVLink.
int32_t qb200_mps_broadcast(qb200_ctx* ctx, qb200_mps** inout, int32_t root) {
    if (!ctx || !inout) QB_FAIL(ctx, QB200_E_INVALID, "mps_broadcast: null argument");
    if (!ctx->nccl_comm) QB_FAIL(ctx, QB200_E_COMM, "communicator not initialised");
    const bool sender = (ctx->comm_rank == root);
    if (sender && !*inout) QB_FAIL(ctx, QB200_E_INVALID, "mps_broadcast: the root rank must pass its chain");
    if (!sender && *inout) QB_FAIL(ctx, QB200_E_INVALID, "mps_broadcast: receiving ranks must pass NULL");
    if (sender) QB_TRY(check_complete(ctx, *inout));
    if (sender && (*inout)->wide) QB_FAIL(ctx, QB200_E_INVALID, "mps_broadcast: chain is in use");
    Workspace ws(ctx);
    constexpr int MAXN = 2000;  // header: n, form, then (chil, p, chir, lambda length) per site; fits the pinned page
    int64_t* hdr_dev = ws.get<int64_t>(2 + 4 * MAXN);
    if (!hdr_dev) QB_FAIL(ctx, QB200_E_CUDA, "mps_broadcast: workspace allocation failed");
    int64_t* hdr = reinterpret_cast<int64_t*>(ctx->scratch_host);
    const size_t hbytes = sizeof(int64_t) * (2 + 4 * MAXN);
    if (sender) {
        qb200_mps* m = *inout;
        if (m->n > MAXN) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "mps_broadcast: more than %d sites", MAXN);
        memset(hdr, 0, hbytes);
        hdr[0] = m->n;
        hdr[1] = m->form + 16 * m->dtype;
        for (int s = 0; s < m->n; ++s) {
            hdr[2 + 4 * s] = m->chil[s];
            hdr[3 + 4 * s] = m->p[s];
            hdr[4 + 4 * s] = m->chir[s];
            hdr[5 + 4 * s] = (s < m->n - 1 && m->lam[s]) ? (int64_t)m->lam_host[s].size() : -1;
        }
        QB_CUDA(ctx, cudaMemcpyAsync(hdr_dev, hdr, hbytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    QB_TRY(qb_comm_broadcast_bytes(ctx, hdr_dev, hbytes, root));
    QB_CUDA(ctx, cudaMemcpyAsync(hdr, hdr_dev, hbytes, cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    qb200_mps* m = *inout;
    if (!sender) {
        QB_TRY(qb200_mps_create(ctx, (int32_t)hdr[0], &m));
        m->form = (int)(hdr[1] % 16);
        m->dtype = (int)(hdr[1] / 16);
        for (int s = 0; s < m->n; ++s) {
            const int64_t cl = hdr[2 + 4 * s], p = hdr[3 + 4 * s], cr = hdr[4 + 4 * s], ll = hdr[5 + 4 * s];
            c128* d = dev_alloc(ctx, (int64_t)((cl * p * cr * site_esz(m) + sizeof(c128) - 1) / sizeof(c128)));
            if (!d) {
                qb200_mps_free(ctx, m);
                QB_FAIL(ctx, QB200_E_CUDA, "mps_broadcast: out of device memory");
            }
            set_site_dev(ctx, m, s, d, cl, p, cr);
            if (s < m->n - 1 && ll >= 0) {
                void* l = nullptr;
                if (cudaMallocAsync(&l, sizeof(double) * (size_t)std::max<int64_t>(ll, 1), ctx->stream) != cudaSuccess) {
                    qb200_mps_free(ctx, m);
                    QB_FAIL(ctx, QB200_E_CUDA, "mps_broadcast: out of device memory");
                }
                m->lam[s] = (double*)l;
                m->lam_host[s].assign((size_t)ll, 0.0);
            }
        }
    }
    int32_t r = qb_comm_group(ctx, true);  // all tensors of the chain in ONE fused NCCL launch
    for (int s = 0; s < m->n && r == QB200_OK; ++s) {
        r = qb_comm_broadcast_bytes(ctx, m->site[s], site_esz(m) * (size_t)(m->chil[s] * m->p[s] * m->chir[s]), root);
        if (r == QB200_OK && s < m->n - 1 && m->lam[s])
            r = qb_comm_broadcast_bytes(ctx, m->lam[s], sizeof(double) * m->lam_host[s].size(), root);
    }
    {
        int32_t r2 = qb_comm_group(ctx, false);
        if (r == QB200_OK) r = r2;
    }
    if (r == QB200_OK && !sender)  // host mirrors of the Schmidt vectors (truncate! reads them element-wise on the host)
        for (int s = 0; s < m->n - 1 && r == QB200_OK; ++s)
            if (m->lam[s] && cudaMemcpyAsync(m->lam_host[s].data(), m->lam[s], sizeof(double) * m->lam_host[s].size(),
                                             cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess)
                r = QB200_E_CUDA;
    if (r == QB200_OK && qb_stream_sync(ctx) != cudaSuccess) r = QB200_E_CUDA;
    if (r != QB200_OK) {
        if (!sender) qb200_mps_free(ctx, m);
        return r;
    }
    *inout = m;
    return QB200_OK;
}

// expect(ψ, observables) (Chain.jl:724-735): the reference's own composition on the device
int32_t qb200_mps_expect(qb200_ctx* ctx, const qb200_mps* m, int32_t nobs, const int32_t* nlanes, const int32_t* sites,
                         const void* ops, double result[2]) {
    QB_TRY(check_complete(ctx, m));
    QB_WIDE(ctx, m, false);
    if (nobs < 0 || !result || (nobs > 0 && (!nlanes || !sites || !ops))) QB_FAIL(ctx, QB200_E_INVALID, "expect: bad argument");
    qb200_mps* phi = nullptr;
    QB_TRY(qb200_mps_copy(ctx, m, &phi));
    const c128* src = (const c128*)ops;
    int32_t r = QB200_OK;
    for (int i = 0; i < nobs && r == QB200_OK; ++i) {
        const int s = sites[i];
        if (nlanes[i] == 1) {
            if (s < 0 || s >= m->n) { ctx->err = "expect: site out of range"; r = QB200_E_INVALID; break; }
            r = qb200_mps_evolve1(ctx, phi, s, src);
            src += m->p[s] * m->p[s];
        } else if (nlanes[i] == 2) {
            if (s < 0 || s >= m->n - 1) { ctx->err = "expect: bond out of range"; r = QB200_E_INVALID; break; }
            r = evolve2_core(ctx, phi, s, src, 0, -1.0, 0, 0, nullptr, nullptr, false);
            const int64_t pp = m->p[s] * m->p[s + 1];
            src += pp * pp;
        } else {
            ctx->err = "Invalid number of lanes, maximum is 2";  // Chain.jl:580
            r = QB200_E_INVALID;
        }
    }
    if (r == QB200_OK) r = qb200_mps_overlap(ctx, phi, m, result);
    qb200_mps_free(ctx, phi);
    return r;
}

}  // extern "C"
