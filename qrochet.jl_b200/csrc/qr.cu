// K4: thin QR of a ComplexF64 matrix (LinearAlgebra.qr(::Tensor; left_inds, right_inds, virtualind),
// Chain.jl:367), hand-written for sm_100a.
//
// Panels of 32 columns.  Each panel is factorised by a communication-avoiding TSQR: the rows are cut in
// chunks of <= 192; one CTA per chunk runs an unblocked Householder QR in shared memory (column norms and
// reflector applications by warp-shuffle reductions), keeps its explicit Q_i in place and sends its
// 32 x 32 R_i to a stack that is factorised recursively; a second kernel multiplies the Q_i by the
// matching 32 x 32 block of the upper level's Q.  The trailing matrix is updated with two DMMA GEMMs
// (C = Q_p^H T with split-K, T -= Q_p C): a right-looking block Gram-Schmidt.  The whole pass is run
// twice (A = Q1 R1, Q1 = Q R2, R = R2 R1) -- "twice is enough" -- which restores orthogonality to
// machine precision; Q is produced explicitly, so there is no separate "form Q" phase.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "mma.cuh"

using namespace qb;

namespace {

constexpr int QW = 32;    // panel width
constexpr int QCH = 192;  // max rows per TSQR leaf
constexpr int QPITCH = QCH + 1;
constexpr int QR_THREADS = 256;
constexpr size_t LEAF_SMEM = (size_t)(QW * QPITCH + QW * QW + 4 * QW) * sizeof(c128);

__device__ __forceinline__ void chunk_range(int rows, int nch, int c, int& begin, int& cnt) {
    int base = rows / nch, rem = rows % nch;
    begin = c * base + min(c, rem);
    cnt = base + (c < rem ? 1 : 0);
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < QR_THREADS / 32; ++i) s += red[i];
    return s;
}

// Householder QR of one row chunk (rows x w, rows >= w) in shared memory; Q explicit in place, R to Rdst
__global__ void __launch_bounds__(QR_THREADS, 1)
    tsqr_leaf_kernel(c128* __restrict__ P, int64_t ld, int rows_total, int nch, int w, c128* __restrict__ Rdst,
                     int64_t ldr, int stacked) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* T = reinterpret_cast<c128*>(smem_raw);  // [w][QPITCH] column-major tile
    c128* Rs = T + QW * QPITCH;                   // [w][w] saved R
    c128* tau = Rs + QW * QW;                     // [w]
    double* red = reinterpret_cast<double*>(tau + QW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int r0, rows;
    chunk_range(rows_total, nch, blockIdx.x, r0, rows);
    c128* Pc = P + r0;
    for (int e = tid; e < rows * w; e += QR_THREADS) {
        int r = e % rows, c = e / rows;
        T[c * QPITCH + r] = Pc[r + (int64_t)c * ld];
    }
    __syncthreads();

    for (int j = 0; j < w; ++j) {
        // |x(j+1:)|^2
        double part = 0.0;
        for (int r = j + 1 + tid; r < rows; r += QR_THREADS) {
            c128 v = T[j * QPITCH + r];
            part += v.x * v.x + v.y * v.y;
        }
        double xn2 = block_sum(part, red);
        c128 alpha = T[j * QPITCH + j];
        c128 tj = make_double2(0.0, 0.0), scl = make_double2(0.0, 0.0);
        double beta = alpha.x;
        if (xn2 != 0.0 || alpha.y != 0.0) {
            double nrm = sqrt(alpha.x * alpha.x + alpha.y * alpha.y + xn2);
            beta = (alpha.x >= 0.0) ? -nrm : nrm;
            tj = make_double2((beta - alpha.x) / beta, -alpha.y / beta);
            // 1 / (alpha - beta)
            double dr = alpha.x - beta, di = alpha.y, den = dr * dr + di * di;
            scl = make_double2(dr / den, -di / den);
        }
        __syncthreads();
        // v = x / (alpha - beta) below the diagonal, v_j = 1 implicit; R_jj = beta
        for (int r = j + 1 + tid; r < rows; r += QR_THREADS) T[j * QPITCH + r] = cmul(T[j * QPITCH + r], scl);
        if (tid == 0) {
            T[j * QPITCH + j] = make_double2(beta, 0.0);
            tau[j] = tj;
        }
        __syncthreads();
        // A(j:, c) <- (I - conj(tau) v v^H) A(j:, c) for c > j : one warp per column
        for (int c = j + 1 + warp; c < w; c += QR_THREADS / 32) {
            double sr = 0.0, si = 0.0;
            for (int r = j + 1 + lane; r < rows; r += 32) {
                c128 s = cmulc(T[c * QPITCH + r], T[j * QPITCH + r]);  // a * conj(v)
                sr += s.x;
                si += s.y;
            }
            sr = warp_sum(sr);
            si = warp_sum(si);
            c128 ajc = T[c * QPITCH + j];
            c128 s = make_double2(sr + ajc.x, si + ajc.y);       // v^H a (v_j = 1)
            c128 f = cmul(cconj(tj), s);                          // conj(tau) * s
            for (int r = j + 1 + lane; r < rows; r += 32)
                T[c * QPITCH + r] = csub(T[c * QPITCH + r], cmul(f, T[j * QPITCH + r]));
            __syncwarp();  // every lane has read ajc before lane 0 overwrites it
            if (lane == 0) T[c * QPITCH + j] = csub(ajc, f);
        }
        __syncthreads();
    }
    // save R (upper triangle)
    for (int e = tid; e < w * w; e += QR_THREADS) {
        int r = e % w, c = e / w;
        Rs[e] = (r <= c) ? T[c * QPITCH + r] : make_double2(0.0, 0.0);
    }
    __syncthreads();
    // form Q in place (LAPACK zung2r)
    for (int j = w - 1; j >= 0; --j) {
        c128 tj = tau[j];
        for (int c = j + 1 + warp; c < w; c += QR_THREADS / 32) {
            double sr = 0.0, si = 0.0;
            for (int r = j + 1 + lane; r < rows; r += 32) {
                c128 s = cmulc(T[c * QPITCH + r], T[j * QPITCH + r]);
                sr += s.x;
                si += s.y;
            }
            sr = warp_sum(sr);
            si = warp_sum(si);
            c128 ajc = T[c * QPITCH + j];  // row j of column c (0 at this point, kept general)
            c128 s = make_double2(sr + ajc.x, si + ajc.y);
            c128 f = cmul(tj, s);
            for (int r = j + 1 + lane; r < rows; r += 32)
                T[c * QPITCH + r] = csub(T[c * QPITCH + r], cmul(f, T[j * QPITCH + r]));
            __syncwarp();  // every lane has read ajc before lane 0 overwrites it
            if (lane == 0) T[c * QPITCH + j] = csub(ajc, f);
        }
        __syncthreads();
        for (int r = tid; r < rows; r += QR_THREADS) {
            c128 v;
            if (r < j)
                v = make_double2(0.0, 0.0);
            else if (r == j)
                v = make_double2(1.0 - tj.x, -tj.y);
            else {
                c128 x = T[j * QPITCH + r];
                v = make_double2(-(tj.x * x.x - tj.y * x.y), -(tj.x * x.y + tj.y * x.x));
            }
            T[j * QPITCH + r] = v;
        }
        __syncthreads();
    }
    for (int e = tid; e < rows * w; e += QR_THREADS) {
        int r = e % rows, c = e / rows;
        Pc[r + (int64_t)c * ld] = T[c * QPITCH + r];
    }
    c128* rd = stacked ? (Rdst + (int64_t)blockIdx.x * w) : Rdst;
    for (int e = tid; e < w * w; e += QR_THREADS) {
        int r = e % w, c = e / w;
        rd[r + (int64_t)c * ldr] = Rs[e];
    }
}

// P_i <- P_i * B_i with B_i = rows [i*w, (i+1)*w) of the upper-level Q (ldq)
__global__ void __launch_bounds__(QR_THREADS)
    tsqr_apply_kernel(c128* __restrict__ P, int64_t ld, int rows_total, int nch, int w, const c128* __restrict__ Qup,
                      int64_t ldq) {
    __shared__ c128 B[QW * QW];
    int r0, rows;
    chunk_range(rows_total, nch, blockIdx.x, r0, rows);
    for (int e = threadIdx.x; e < w * w; e += QR_THREADS) {
        int r = e % w, c = e / w;
        B[e] = Qup[(int64_t)blockIdx.x * w + r + (int64_t)c * ldq];
    }
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += QR_THREADS) {
        c128 row[QW];
        c128* p = P + r0 + r;
        for (int c = 0; c < w; ++c) row[c] = p[(int64_t)c * ld];
        for (int c2 = 0; c2 < w; ++c2) {
            c128 acc = make_double2(0.0, 0.0);
            for (int c = 0; c < w; ++c) acc = cadd(acc, cmul(row[c], B[c + c2 * w]));
            p[(int64_t)c2 * ld] = acc;
        }
    }
}

// squared column norms, one warp per column
__global__ void colnorm2_kernel(const c128* __restrict__ A, int64_t ld, int64_t m, int64_t ncols, double* __restrict__ out) {
    const int64_t col = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= ncols) return;
    const int lane = threadIdx.x & 31;
    const c128* x = A + col * ld;
    double acc = 0.0;
    for (int64_t r = lane; r < m; r += 32) {
        c128 v = x[r];
        acc += v.x * v.x + v.y * v.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[col] = acc;
}

// a reproducible pseudo-random column (entries in (-1, 1)): the replacement of a numerically dependent column
__global__ void random_column_kernel(c128* __restrict__ x, int64_t m, uint64_t seed) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(2 * i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const double re = (double)(z >> 40) * (1.0 / 8388608.0) - 1.0;
        const double im = (double)((z >> 16) & 0xFFFFFF) * (1.0 / 8388608.0) - 1.0;
        x[i] = make_double2(re, im);
    }
}

// R2 = I + triu(D, 1) + diag(D) / 2 with D = E - I, E = Q1^H Q1 (in place in E): the Cholesky factor of I + D to first
// order in D.  The largest |D_ij| goes to *dmax (bit pattern of a non-negative double orders like an integer).
__global__ void refine_r2_kernel(c128* __restrict__ E, int64_t k, unsigned long long* __restrict__ dmax) {
    double mx = 0.0;
    const int64_t total = k * k;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx % k, j = idx / k;
        c128 v = E[idx];
        if (i == j) {
            mx = fmax(mx, fabs(v.x - 1.0));
            E[idx] = make_double2(0.5 * (1.0 + v.x), 0.0);
        } else {
            mx = fmax(mx, fmax(fabs(v.x), fabs(v.y)));
            if (i > j) E[idx] = make_double2(0.0, 0.0);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx > 0.0) atomicMax(dmax, (unsigned long long)__double_as_longlong(mx));
}

__global__ void zero_kernel(c128* x, int64_t rows, int64_t cols, int64_t ld) {
    int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        x[i % rows + (i / rows) * ld] = make_double2(0.0, 0.0);
}

}  // namespace

static int32_t tsqr(qb200_ctx* ctx, c128* P, int64_t ld, int rows, int w, c128* Rout, int64_t ldr) {
    int nch = (rows + QCH - 1) / QCH;
    if (nch == 1) {
        tsqr_leaf_kernel<<<1, QR_THREADS, LEAF_SMEM, ctx->stream>>>(P, ld, rows, 1, w, Rout, ldr, 0);
        QB_LAUNCH_CHECK(ctx);
        return QB200_OK;
    }
    Workspace ws(ctx);
    int srows = nch * w;
    c128* stack = ws.get<c128>((size_t)srows * w);
    if (!stack) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
    tsqr_leaf_kernel<<<nch, QR_THREADS, LEAF_SMEM, ctx->stream>>>(P, ld, rows, nch, w, stack, srows, 1);
    QB_LAUNCH_CHECK(ctx);
    QB_TRY(tsqr(ctx, stack, srows, srows, w, Rout, ldr));
    tsqr_apply_kernel<<<nch, QR_THREADS, 0, ctx->stream>>>(P, ld, rows, nch, w, stack, srows);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// TSQR + Gram-Schmidt on a panel of width <= 64 made of <= 2 sub-panels of 32 (the robust path)
static int32_t tsqr_panel(qb200_ctx* ctx, int64_t m, int w, c128* P, int64_t ld, c128* R, int64_t ldr) {
    const c128 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), mone = make_double2(-1.0, 0.0);
    for (int j0 = 0; j0 < w; j0 += QW) {
        int ww = std::min(QW, w - j0);
        c128* Pj = P + (int64_t)j0 * ld;
        QB_TRY(tsqr(ctx, Pj, ld, (int)m, ww, R + j0 + (int64_t)j0 * ldr, ldr));
        int nt = w - j0 - ww;
        if (nt > 0) {
            c128* T = P + (int64_t)(j0 + ww) * ld;
            c128* C = R + j0 + (int64_t)(j0 + ww) * ldr;
            QB_TRY(qb_gemm(ctx, 2, 0, ww, nt, m, one, Pj, ld, T, ld, zero, C, ldr));
            QB_TRY(qb_gemm(ctx, 0, 0, m, nt, ww, mone, Pj, ld, C, ldr, one, T, ld));
        }
    }
    return QB200_OK;
}

// Relative size below which what is left of a column after the projection against the earlier columns is certainly
// rounding noise (the projection subtracts ~j0 terms, each carrying eps of the column's norm).  Used for the first panel
// (nothing to re-project against) and as a backstop; the decisive test is the re-projection below.
static inline double dep_tol(int64_t j0, int w) { return 16.0 * 1.1102230246251565e-16 * std::sqrt((double)(j0 + w)); }
// A column that lost more than this fraction of its norm in the projection sends its panel to robust_panel
constexpr double CAREFUL_LOSS = 1e-6;

// Panel factorisation that is safe for rank-deficient input.  A column that is numerically dependent on the earlier
// columns is rounding noise once those have been projected out; normalising it (what any Gram-Schmidt or Householder
// panel step does) yields a "completion" vector that still lies in the span of the earlier columns -- for structurally
// rank-deficient tensors (an MPO-applied state: exact zeros, exact repetitions) entirely so -- and a second Gram-Schmidt
// pass cannot repair that: Q comes out with non-orthogonal columns and every later column projected on it is damaged
// (measured on B200: |Q^H Q - I| = 0.23 on a 5120 x 2560 matrix of rank 1536; canonize! of H|psi> changed the norm of
// the state by 23 %).  The test for "this Q column is garbage" must not depend on an eps-sized threshold (the input of a
// sweep carries the rounding of the previous sites, so what is left of a dependent column is not just projection noise):
// it is Kahan's "twice is enough" criterion.  Factor the panel, project its Q once more on the earlier columns and look at
// what survives: a healthy column keeps its unit norm, a garbage column collapses.  Columns whose norm falls below 1/2
// are replaced by a pseudo-random vector made orthogonal (twice) to ALL earlier columns and the panel is factored again;
// their column of R keeps only the components of the original column along the panel's other columns (nothing along
// the completion vector).  When no column collapses the re-projection is folded into R (R[0:j0, panel] += C2 Rpp) and
// the panel is re-orthonormalised -- a block Gram-Schmidt step with immediate re-orthogonalisation.  This is what
// LAPACK's Householder QR delivers implicitly: orthonormal Q whatever the rank.  `save` holds the panel as it was before
// the factorisation (m x w, ld = m) and is updated with the replacements; Rtop = R(0, j0): the rows of R above the panel.
static int32_t robust_panel(qb200_ctx* ctx, int64_t m, int w, c128* P, int64_t ld, c128* Rpp, int64_t ldr, const c128* Qall,
                            int64_t ldq, int64_t j0, const double* before2_dev, c128* save) {
    const c128 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), mone = make_double2(-1.0, 0.0);
    std::vector<char> replaced(w, 0);
    Workspace ws(ctx);
    c128* coef = nullptr;
    c128* orig = nullptr;  // the dependent columns as they were (their components along the panel's other columns go to R)
    c128* xd = nullptr;    // their replacements, gathered (m x #dependent)
    c128* C2 = (j0 > 0) ? ws.get<c128>((size_t)(j0 * w)) : nullptr;
    c128* R2 = ws.get<c128>((size_t)(w * w));
    c128* Rt = ws.get<c128>((size_t)(w * w));
    double* nu2 = ws.get<double>((size_t)w);
    if ((j0 > 0 && !C2) || !R2 || !Rt || !nu2) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
    c128* Rtop = Rpp - j0;  // same columns of R, rows 0 .. j0-1
    const double tol = dep_tol(j0, w);
    // every attempt that does not succeed enlarges the set of replaced columns, so w + 1 attempts always suffice
    for (int attempt = 0; attempt <= w; ++attempt) {
        QB_TRY(tsqr_panel(ctx, m, w, P, ld, Rpp, ldr));
        for (int c = 0; c < w; ++c)  // a completion vector carries nothing of the original column
            if (replaced[c]) QB_CUDA(ctx, cudaMemsetAsync(Rpp + (int64_t)c * ldr, 0, sizeof(c128) * w, ctx->stream));
        double* host = ctx->scratch_host;
        if (attempt == 0) {  // the common case: no column lost more than CAREFUL_LOSS of its norm -> the panel is healthy
            QB_CUDA(ctx, cudaMemcpy2DAsync(host, sizeof(c128), Rpp, (size_t)(ldr + 1) * sizeof(c128), sizeof(c128), (size_t)w,
                                           cudaMemcpyDeviceToHost, ctx->stream));
            QB_CUDA(ctx, cudaMemcpyAsync(host + 2 * w, before2_dev, sizeof(double) * w, cudaMemcpyDeviceToHost, ctx->stream));
            QB_CUDA(ctx, qb_stream_sync(ctx));
            bool healthy = true;
            for (int c = 0; c < w; ++c) {
                const double d2 = host[2 * c] * host[2 * c] + host[2 * c + 1] * host[2 * c + 1];
                if (!(d2 > CAREFUL_LOSS * CAREFUL_LOSS * host[2 * w + c])) healthy = false;
            }
            if (healthy) return QB200_OK;
        }
        if (j0 > 0) {  // what survives of the panel's Q when the earlier columns are projected out once more
            QB_TRY(qb_gemm(ctx, 2, 0, j0, w, m, one, Qall, ldq, P, ld, zero, C2, j0));
            QB_TRY(qb_gemm(ctx, 0, 0, m, w, j0, mone, Qall, ldq, C2, j0, one, P, ld));
            colnorm2_kernel<<<(unsigned)((w + 7) / 8), 256, 0, ctx->stream>>>(P, ld, m, w, nu2);
            QB_LAUNCH_CHECK(ctx);
        }
        // diag(R) of the panel, the squared norms at the start of the pass and the survival norms -> host (one sync)
        QB_CUDA(ctx, cudaMemcpy2DAsync(host, sizeof(c128), Rpp, (size_t)(ldr + 1) * sizeof(c128), sizeof(c128), (size_t)w,
                                       cudaMemcpyDeviceToHost, ctx->stream));
        QB_CUDA(ctx, cudaMemcpyAsync(host + 2 * w, before2_dev, sizeof(double) * w, cudaMemcpyDeviceToHost, ctx->stream));
        if (j0 > 0) QB_CUDA(ctx, cudaMemcpyAsync(host + 3 * w, nu2, sizeof(double) * w, cudaMemcpyDeviceToHost, ctx->stream));
        QB_CUDA(ctx, qb_stream_sync(ctx));
        std::vector<int> dep;
        for (int c = 0; c < w; ++c) {
            if (replaced[c]) continue;
            const double d2 = host[2 * c] * host[2 * c] + host[2 * c + 1] * host[2 * c + 1];
            const bool noise = d2 <= tol * tol * host[2 * w + c];
            const bool collapsed = j0 > 0 && host[3 * w + c] < 0.25;
            if (noise || collapsed) dep.push_back(c);
        }
        if (dep.empty()) {
            if (j0 > 0) {  // fold the re-projection into R and re-orthonormalise the panel: Q_p = Q' R2, Rpp <- R2 Rpp
                QB_TRY(qb_gemm(ctx, 0, 0, j0, w, w, one, C2, j0, Rpp, ldr, one, Rtop, ldr));
                QB_CUDA(ctx, cudaMemsetAsync(R2, 0, sizeof(c128) * (size_t)(w * w), ctx->stream));  // tsqr_panel writes the upper blocks only
                QB_TRY(tsqr_panel(ctx, m, w, P, ld, R2, w));
                QB_TRY(qb_gemm(ctx, 0, 0, w, w, w, one, R2, w, Rpp, ldr, zero, Rt, w));
                QB_TRY(qb_copy_matrix(ctx, w, w, Rt, w, Rpp, ldr, 0));
            }
            break;
        }
        if (attempt == w) QB_FAIL(ctx, QB200_E_NOCONVERGE, "qr: could not complete a rank-deficient panel");
        if (!orig) orig = ws.get<c128>((size_t)(m * w));
        if (!orig) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
        // all dependent columns of the panel at once: random vectors, projected twice on the earlier columns (two GEMM
        // pairs per attempt, not per column)
        if (!xd) xd = ws.get<c128>((size_t)(m * w));
        if (!xd) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
        const int nd = (int)dep.size();
        for (int i = 0; i < nd; ++i) {
            const int c = dep[i];
            QB_CUDA(ctx, cudaMemcpyAsync(orig + (int64_t)c * m, save + (int64_t)c * m, sizeof(c128) * (size_t)m,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
            random_column_kernel<<<(unsigned)std::min<int64_t>((m + 255) / 256, 1024), 256, 0, ctx->stream>>>(
                xd + (int64_t)i * m, m, 0x5EEDull + 1315423911ull * (uint64_t)(j0 + c));
            QB_LAUNCH_CHECK(ctx);
            replaced[c] = 1;
        }
        if (j0 > 0) {
            if (!coef) coef = ws.get<c128>((size_t)(j0 * w));
            if (!coef) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
            for (int rep = 0; rep < 2; ++rep) {
                QB_TRY(qb_gemm(ctx, 2, 0, j0, nd, m, one, Qall, ldq, xd, m, zero, coef, j0));
                QB_TRY(qb_gemm(ctx, 0, 0, m, nd, j0, mone, Qall, ldq, coef, j0, one, xd, m));
            }
        }
        for (int i = 0; i < nd; ++i)
            QB_CUDA(ctx, cudaMemcpyAsync(save + (int64_t)dep[i] * m, xd + (int64_t)i * m, sizeof(c128) * (size_t)m,
                                         cudaMemcpyDeviceToDevice, ctx->stream));
        QB_TRY(qb_copy_matrix(ctx, m, w, save, m, P, ld, 0));
    }
    // column c of the panel's R block: the components of the ORIGINAL column along the panel's earlier columns (a
    // dependency inside the panel), nothing on or below the diagonal (no component along its completion vector)
    bool any = false;
    for (int c = 0; c < w; ++c) any = any || replaced[c];
    if (any) {
        // one GEMM for all of them (columns of `orig` that were never saved hold junk; their results are not used)
        QB_TRY(qb_gemm(ctx, 2, 0, w, w, m, one, P, ld, orig, m, zero, Rt, w));
        for (int c = 0; c < w; ++c)
            if (replaced[c]) {
                ctx->qr_last_dependent++;
                QB_CUDA(ctx, cudaMemsetAsync(Rpp + (int64_t)c * ldr, 0, sizeof(c128) * (size_t)w, ctx->stream));
                if (c > 0)
                    QB_CUDA(ctx, cudaMemcpyAsync(Rpp + (int64_t)c * ldr, Rt + (int64_t)c * w, sizeof(c128) * (size_t)c,
                                                 cudaMemcpyDeviceToDevice, ctx->stream));
            }
    }
    return QB200_OK;
}

// one right-looking block Gram-Schmidt pass: Q (m x k, in place) = Q' R, R k x k upper triangular (zeroed first).
// Panels of 64 columns.  Fast path (m % 64 == 0, full panels): Cholesky-QR2 on the panel with the Jacobi gram /
// update kernels (Gram 64 x 64 -> scaled Cholesky in shared memory -> P <- P R^-1, twice); when a scaled pivot says
// the panel is too ill conditioned for a Gram-based step, or a column has lost more than CAREFUL_LOSS of its norm to the
// earlier columns, the saved panel is restored and factorised by robust_panel (Householder TSQR, re-projection,
// completion of dependent columns) instead.  Trailing updates: C = Q_p^H T (split-K GEMM), T -= Q_p C.
// optimistic != nullptr: every panel takes the Cholesky-QR fast path WITHOUT the per-panel read-back of its failure flag
// (one host round trip per 64 columns: 32 per bulk TEBD bond, each a sleep / wake-up of the worker thread -- what made
// the sweep time depend on how busy the host's cores are).  The flag accumulates on the device and is read ONCE at the
// end; *optimistic = true means "some panel failed, the contents of Q and R are garbage": the caller restores its input
// and runs the careful pass.  Only possible when every panel is a full fast-path panel.
static int32_t bgs_pass(qb200_ctx* ctx, int64_t m, int64_t k, c128* Q, int64_t ldq, c128* R, int64_t ldr,
                        bool* optimistic = nullptr) {
    const c128 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), mone = make_double2(-1.0, 0.0);
    zero_kernel<<<(unsigned)std::min<int64_t>((k * k + 255) / 256, 4096), 256, 0, ctx->stream>>>(R, k, k, ldr);
    QB_LAUNCH_CHECK(ctx);
    const int PW = 64;
    const bool fast_ok = (m % 64 == 0) && m >= 64 && !getenv("QB200_NO_CHOLQR");
    const bool opt = optimistic && fast_ok && (k % 64 == 0);
    if (optimistic) *optimistic = false;
    Workspace ws(ctx);
    c128 *Gpart = nullptr, *Wbuf = nullptr, *R1 = nullptr, *R2 = nullptr;
    int* flags = nullptr;
    c128* save = ws.get<c128>((size_t)(m * PW));
    double* before2 = ws.get<double>((size_t)k);  // squared column norms at the start of the pass
    if (!save || !before2) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
    colnorm2_kernel<<<(unsigned)((k + 7) / 8), 256, 0, ctx->stream>>>(Q, ldq, m, k, before2);
    QB_LAUNCH_CHECK(ctx);
    if (fast_ok) {
        Gpart = ws.get<c128>(qb_cholqr_gpart_elems(ctx));
        Wbuf = ws.get<c128>(64 * 64);
        R1 = ws.get<c128>(64 * 64);
        R2 = ws.get<c128>(64 * 64);
        flags = ws.get<int>(2);
        if (!Gpart || !Wbuf || !R1 || !R2 || !flags) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
        if (opt) QB_CUDA(ctx, cudaMemsetAsync(flags, 0, 2 * sizeof(int), ctx->stream));
    }
    // super-panels of 128 columns = two Cholesky-QR sub-panels of 64; the trailing matrix is updated once per
    // super-panel (K = 128 GEMMs: half the launches, twice the depth)
    const int SW = 128;
    for (int64_t s0 = 0; s0 < k; s0 += SW) {
        int sw = (int)std::min<int64_t>(SW, k - s0);
        for (int64_t j0 = s0; j0 < s0 + sw; j0 += PW) {
            int w = (int)std::min<int64_t>(PW, s0 + sw - j0);
            c128* P = Q + j0 * ldq;
            c128* Rpp = R + j0 + j0 * ldr;
            bool done = false;
            if (opt) {
                QB_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int), ctx->stream));  // flags[1] (failure) accumulates
                QB_TRY(qb_cholqr_panel_step(ctx, P, ldq, m, R1, 64, Gpart, Wbuf, flags, flags + 1, before2 + j0,
                                            CAREFUL_LOSS));
                QB_TRY(qb_cholqr_panel_step(ctx, P, ldq, m, R2, 64, Gpart, Wbuf, flags, flags + 1, nullptr, 0.0));
                QB_TRY(qb_gemm(ctx, 0, 0, PW, PW, PW, one, R2, 64, R1, 64, zero, Rpp, ldr));
                done = true;
            } else {
                QB_TRY(qb_copy_matrix(ctx, m, w, P, ldq, save, m, 0));
            }
            if (!opt && fast_ok && w == PW) {
                QB_CUDA(ctx, cudaMemsetAsync(flags, 0, 2 * sizeof(int), ctx->stream));
                QB_TRY(qb_cholqr_panel_step(ctx, P, ldq, m, R1, 64, Gpart, Wbuf, flags, flags + 1, before2 + j0,
                                            CAREFUL_LOSS));
                QB_TRY(qb_cholqr_panel_step(ctx, P, ldq, m, R2, 64, Gpart, Wbuf, flags, flags + 1, nullptr, 0.0));
                QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
                QB_CUDA(ctx, qb_stream_sync(ctx));
                int failed = *reinterpret_cast<int*>(ctx->scratch_host);
                if (!failed) {
                    QB_TRY(qb_gemm(ctx, 0, 0, PW, PW, PW, one, R2, 64, R1, 64, zero, Rpp, ldr));
                    done = true;
                } else {
                    QB_TRY(qb_copy_matrix(ctx, m, PW, save, m, P, ldq, 0));
                }
            }
            if (!done) QB_TRY(robust_panel(ctx, m, w, P, ldq, Rpp, ldr, Q, ldq, j0, before2 + j0, save));
            int64_t nin = s0 + sw - j0 - w;  // rest of the super-panel
            if (nin > 0) {
                c128* T = Q + (j0 + w) * ldq;
                c128* C = R + j0 + (j0 + w) * ldr;
                QB_TRY(qb_gemm(ctx, 2, 0, w, nin, m, one, P, ldq, T, ldq, zero, C, ldr));
                QB_TRY(qb_gemm(ctx, 0, 0, m, nin, w, mone, P, ldq, C, ldr, one, T, ldq));
            }
        }
        int64_t nt = k - s0 - sw;
        if (nt > 0) {
            c128* P = Q + s0 * ldq;
            c128* T = Q + (s0 + sw) * ldq;
            c128* C = R + s0 + (s0 + sw) * ldr;
            QB_TRY(qb_gemm(ctx, 2, 0, sw, nt, m, one, P, ldq, T, ldq, zero, C, ldr));   // C = P^H T
            QB_TRY(qb_gemm(ctx, 0, 0, m, nt, sw, mone, P, ldq, C, ldr, one, T, ldq));   // T -= P C
        }
    }
    if (opt) {
        QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        QB_CUDA(ctx, qb_stream_sync(ctx));
        *optimistic = *reinterpret_cast<int*>(ctx->scratch_host) != 0;
    }
    return QB200_OK;
}

// first pass over a fresh copy of A: optimistic, with the careful pass as the fallback
static int32_t first_pass(qb200_ctx* ctx, int64_t m, int64_t k, const c128* A, int64_t lda, c128* Q, int64_t ldq, c128* R,
                          int64_t ldr) {
    static const bool optimistic_on = [] {  // QB200_QR_OPTIMISTIC=0: per-panel read-back as before (A/B switch)
        const char* e = getenv("QB200_QR_OPTIMISTIC");
        return !(e && e[0] == '0');
    }();
    if (optimistic_on) {
        bool failed = false;
        QB_TRY(bgs_pass(ctx, m, k, Q, ldq, R, ldr, &failed));
        if (!failed) return QB200_OK;
        QB_TRY(qb_copy_matrix(ctx, m, k, A, lda, Q, ldq, 0));
    }
    return bgs_pass(ctx, m, k, Q, ldq, R, ldr);
}

int32_t qb_qr_init(qb200_ctx* ctx) {
    QB_CUDA(ctx, cudaFuncSetAttribute(tsqr_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LEAF_SMEM));
    return QB200_OK;
}

int32_t qb_qr_matrix(qb200_ctx* ctx, int64_t m, int64_t n, const c128* A, int64_t lda, c128* Q, int64_t ldq,
                     c128* R, int64_t ldr, int passes) {
    if (m <= 0 || n <= 0) QB_FAIL(ctx, QB200_E_INVALID, "qr: empty matrix");
    if (m > INT32_MAX / 2) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "qr: too many rows");
    const int64_t k = std::min(m, n);
    const c128 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
    ctx->qr_last_dependent = 0;
    QB_TRY(qb_copy_matrix(ctx, m, k, A, lda, Q, ldq, 0));
    Workspace ws(ctx);
    c128* R1 = ws.get<c128>((size_t)k * k);
    c128* R2 = ws.get<c128>((size_t)k * k);
    if (!R1 || !R2) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
    static const bool refine_enabled = [] {
        const char* e = getenv("QB200_QR_REFINE");
        return !(e && e[0] == '0');
    }();
    bool refined = false;
    if (passes == QB_QR_R_ONLY && refine_enabled) {
        // Only R is wanted (the SVD preconditioner): the second Gram-Schmidt pass exists to remove the loss of
        // orthogonality D = Q1^H Q1 - I ~ kappa eps of the first from R.  To first order the Cholesky factor of I + D
        // is R2 = I + triu(D, 1) + diag(D)/2, so R = R2 R1 costs two GEMMs (k x k x m and k^3) instead of a second
        // latency-bound pass over 2k/64 panels; the neglected term is O(|D|^2), so the shortcut is taken only when
        // max |D_ij| <= 1e-9 (and the first pass met no dependent column); otherwise the full second pass runs.
        QB_TRY(first_pass(ctx, m, k, A, lda, Q, ldq, R1, k));
        if (ctx->qr_last_dependent == 0) {
            unsigned long long* dmax = ws.get<unsigned long long>(1);
            if (!dmax) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
            QB_CUDA(ctx, cudaMemsetAsync(dmax, 0, sizeof(unsigned long long), ctx->stream));
            QB_TRY(qb_gemm(ctx, 2, 0, k, k, m, one, Q, ldq, Q, ldq, zero, R2, k));
            refine_r2_kernel<<<(unsigned)std::min<int64_t>((k * k + 255) / 256, 2048), 256, 0, ctx->stream>>>(R2, k, dmax);
            QB_LAUNCH_CHECK(ctx);
            QB_TRY(qb_gemm(ctx, 0, 0, k, k, k, one, R2, k, R1, k, zero, R, ldr));
            QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, dmax, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            QB_CUDA(ctx, qb_stream_sync(ctx));
            refined = ctx->scratch_host[0] <= 1e-9;
        }
        if (!refined) {
            QB_TRY(bgs_pass(ctx, m, k, Q, ldq, R2, k));
            QB_TRY(qb_gemm(ctx, 0, 0, k, k, k, one, R2, k, R1, k, zero, R, ldr));
        }
    } else if (passes <= 1 && passes != QB_QR_R_ONLY) {
        // one pass: R is backward stable (as for modified Gram-Schmidt), Q is orthonormal only to kappa(A) eps
        QB_TRY(first_pass(ctx, m, k, A, lda, Q, ldq, R1, k));
        QB_TRY(qb_copy_matrix(ctx, k, k, R1, k, R, ldr, 0));
    } else {
        QB_TRY(first_pass(ctx, m, k, A, lda, Q, ldq, R1, k));
        QB_TRY(bgs_pass(ctx, m, k, Q, ldq, R2, k));
        QB_TRY(qb_gemm(ctx, 0, 0, k, k, k, one, R2, k, R1, k, zero, R, ldr));
    }
    if (n > k) QB_TRY(qb_gemm(ctx, 2, 0, k, n - k, m, one, Q, ldq, A + k * lda, lda, zero, R + k * ldr, ldr));
    return QB200_OK;
}

extern "C" int32_t qb200_qr(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* order, int32_t nleft,
                            qb200_tensor* Q, qb200_tensor* R) {
    if (!ctx || !A || !order || !Q || !R) QB_FAIL(ctx, QB200_E_INVALID, "qr: null argument");
    Workspace ws(ctx);
    const c128* mat;
    int64_t m, n;
    QB_TRY(qb_matricize(ctx, A, order, nleft, ws, &mat, &m, &n));
    int64_t k = std::min(m, n);
    if (Q->dtype != A->dtype || R->dtype != A->dtype || Q->numel() != m * k || R->numel() != k * n)
        QB_FAIL(ctx, QB200_E_INVALID, "qr: Q must hold rows*k and R k*cols entries of the type of A");
    if (A->dtype == QB200_C64) {  // factorised in FP64, factors narrowed to float2
        c128* Qw = ws.get<c128>((size_t)std::max<int64_t>(m * k, 1));
        c128* Rw = ws.get<c128>((size_t)std::max<int64_t>(k * n, 1));
        if (!Qw || !Rw) QB_FAIL(ctx, QB200_E_CUDA, "qr: workspace allocation failed");
        QB_TRY(qb_qr_matrix(ctx, m, n, mat, m, Qw, m, Rw, k, 2));
        QB_TRY(qb_narrow_c128(ctx, Qw, Q->data, m * k));
        return qb_narrow_c128(ctx, Rw, R->data, k * n);
    }
    return qb_qr_matrix(ctx, m, n, mat, m, (c128*)Q->data, m, (c128*)R->data, k, 2);
}
