// K5: QR-preconditioned one-sided block-Jacobi SVD of a ComplexF64 matrix, hand-written for sm_100a.
//
// A (m x n), k = min(m, n): B = A (m >= n) or A^H, columns sorted by norm (B0 = B P), B0 = Q R by K4 (qr.cu; only
// R is kept), and the Jacobi iteration runs on X = R^H (k x k, zero-padded to multiples of 64) -- Drmac-Veselic
// preconditioning.  Columns are grouped in blocks of 32; a step of the round-robin tournament pairs the
// nb = np/32 blocks into nb/2 disjoint pairs and runs three kernels over all pairs at once:
//   gram   : G_p = P_p^H P_p (64 x 64) for the panel P_p = [X_I X_J] -- the full Hermitian-upper Gram at step 0
//            of a sweep, only the 32 x 32 cross block X_I^H X_J afterwards (the diagonal blocks are carried by
//            the evd kernel) -- persistent grid over (pair, 32-row chunk) items, DMMA tiles, 3-stage cp.async
//            pipeline, partial sums written per CTA;
//   evd    : one CTA (17 warps) per pair sums the partials in a fixed order (deterministic), runs a two-sided cyclic
//            Jacobi on the 64 x 64 Hermitian G in shared memory (32 disjoint rotations per parallel step)
//            and leaves the product of the rotations, the unitary W_p; pairs that are already orthogonal
//            are flagged and skipped by the update;
//   update : X_p <- X_p W_p in place with DMMA tiles, the W slice of each warp in registers, complex products by the
//            3M (Gauss) scheme: 3 real DMMA + operand sums per complex tile instead of 4 (also in the cross Gram).
// nb-1 steps make a sweep; sweeps repeat until the largest |x_i^H x_j| / (|x_i||x_j|) seen in a sweep is below
// tolerance.  sigma_j = |x_j|.  The rotations are NOT accumulated: with R^H = Xn S Vs^H the right factor is
// V = P Xn and the left factor comes from ONE GEMM, Y = B0 Xn S^-1 (qb_svd_emit).  The rotations only ever come
// from Gram entries of the columns they are applied to, and W_p is a product of exact plane rotations, so
// singular values keep norm-wise accuracy ~ eps * sigma_1 (same class as LAPACK gesdd).
#include <algorithm>
#include <cmath>
#include <numeric>
#include <cstdlib>

#include "common.cuh"
#include "mma.cuh"
#include "jacobi_rr.cuh"

using namespace qb;

namespace {

constexpr int GLD = 65;  // shared-memory pitch of the 64 x 64 matrices in the evd kernel

// ---------------------------------------------------------------------------------------------
// gram kernel: G_p = P_p^H P_p for every pair of the step.
//
// Work items are (pair, 32-row chunk), flattened pair-major and cut into equal contiguous ranges, one per CTA of
// a persistent grid of 2 x #SM CTAs (no tail wave: 2048 items over 296 CTAs at chi = 1024).  A CTA accumulates
// over its items and flushes a partial Gram whenever the pair changes (at most twice: a range is shorter than a
// pair); the evd kernel sums the partials of its pair in CTA order (deterministic).
// Only the 36 upper 8 x 8 tiles of the 8 x 8 tile grid are computed (G is Hermitian) -- dealt to the 8 warps as
// <= 5 tiles each by a compile-time table -- and mirrored on the way out: 5/8 of the DMMA work of the full product.
constexpr int G_BKR = 32, G_NST = 3, G_PITCH = G_BKR + 4;
constexpr size_t GRAM_SMEM = (size_t)G_NST * JP * G_PITCH * sizeof(c128);

struct GramTiles {
    int n;
    int r[5], c[5];
};
__host__ __device__ constexpr GramTiles gram_tiles(int w) {
    switch (w) {
        case 0: return {5, {0, 0, 0, 0, 0}, {0, 1, 2, 3, 4}};
        case 1: return {5, {0, 0, 0, 1, 1}, {5, 6, 7, 1, 2}};
        case 2: return {5, {1, 1, 1, 1, 1}, {3, 4, 5, 6, 7}};
        case 3: return {5, {2, 2, 2, 2, 2}, {2, 3, 4, 5, 6}};
        case 4: return {5, {2, 3, 3, 3, 3}, {7, 3, 4, 5, 6}};
        case 5: return {5, {3, 4, 4, 4, 4}, {7, 4, 5, 6, 7}};
        case 6: return {5, {5, 5, 5, 6, 6}, {5, 6, 7, 6, 7}};
        default: return {1, {7, 0, 0, 0, 0}, {7, 0, 0, 0, 0}};
    }
}

__device__ __forceinline__ void gram_item_range(int cta, int ncta, int total, int& lo, int& hi) {
    lo = (int)(((long long)cta * total) / ncta);
    hi = (int)(((long long)(cta + 1) * total) / ncta);
}

// cross block only (rows of block I x columns of block J): 16 tiles, 2 per warp
__host__ __device__ constexpr GramTiles gram_tiles_cross(int w) {
    return {2, {w >> 1, w >> 1, 0, 0, 0}, {4 + 2 * (w & 1), 5 + 2 * (w & 1), 0, 0, 0}};
}
template <int W, int CROSS>
__host__ __device__ constexpr GramTiles gram_tile_table() {
    return CROSS ? gram_tiles_cross(W) : gram_tiles(W);
}

// The cross-block instantiation (CROSS = 1, all but one step of a sweep) uses the 3M complex product:
// conj(a) b = (P + Q) + i (S - P + Q) with P = ar br, Q = ai bi, S = (ar - ai)(br + bi): 3 DMMA + operand sums per
// complex 8x8x4 tile instead of 4.  cr / ci / cs then hold P / Q / S until the flush.
template <int W, int CROSS>
__device__ __forceinline__ void gram_mma_chunk(const c128* __restrict__ ps, int g, int t, double (&cr)[5][2],
                                               double (&ci)[5][2], double (&cs)[2][2]) {
    constexpr GramTiles T = gram_tile_table<W, CROSS>();
#pragma unroll
    for (int kk = 0; kk < G_BKR / 4; ++kk) {
        if constexpr (CROSS) {
            // both tiles of a warp share the row tile: one A fragment, two B fragments
            const c128 a = ps[(T.r[0] * 8 + g) * G_PITCH + t + kk * 4];
            const double as = a.x - a.y;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const c128 b = ps[(T.c[i] * 8 + g) * G_PITCH + t + kk * 4];
                dmma884(cr[i], a.x, b.x);
                dmma884(ci[i], a.y, b.y);
                dmma884(cs[i], as, b.x + b.y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < T.n; ++i) {
                c128 a = ps[(T.r[i] * 8 + g) * G_PITCH + t + kk * 4];
                c128 b = ps[(T.c[i] * 8 + g) * G_PITCH + t + kk * 4];
                // G = P^H P: A operand is conj(P)
                dmma884(cr[i], a.x, b.x);
                dmma884(ci[i], a.x, b.y);
                dmma884(cr[i], a.y, b.y);
                dmma884(ci[i], -a.y, b.x);
            }
        }
    }
}

template <int W, int CROSS>
__device__ __forceinline__ void gram_flush(c128* __restrict__ out, int g, int t, double (&cr)[5][2],
                                           double (&ci)[5][2], double (&cs)[2][2]) {
    constexpr GramTiles T = gram_tile_table<W, CROSS>();
#pragma unroll
    for (int i = 0; i < T.n; ++i) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            int row = T.r[i] * 8 + g, col = T.c[i] * 8 + 2 * t + h;
            c128 v;
            if constexpr (CROSS) {
                v = make_double2(cr[i][h] + ci[i][h], cs[i][h] - cr[i][h] + ci[i][h]);
                cs[i][h] = 0.0;
            } else {
                v = make_double2(cr[i][h], ci[i][h]);
            }
            out[row + JP * col] = v;
            if (T.r[i] != T.c[i]) out[col + JP * row] = make_double2(v.x, -v.y);
            cr[i][h] = 0.0;
            ci[i][h] = 0.0;
        }
    }
}

template <int CROSS>
__global__ void __launch_bounds__(256, 2)
    jacobi_gram_kernel(const c128* __restrict__ Z, int64_t ldz, int mp, int nb, int step, int npairs,
                       c128* __restrict__ Gpart) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* Ps = reinterpret_cast<c128*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int nchunk = mp / G_BKR, total = npairs * nchunk;
    int lo, hi;
    gram_item_range(blockIdx.x, gridDim.x, total, lo, hi);
    const int nitems = hi - lo;

    // prefetch cursor: (pair, chunk, I, J) carried incrementally, no division per item
    const int l_row = tid & 31, l_col0 = tid >> 5;
    const int first_pair = (nitems > 0) ? lo / nchunk : 0;
    int pf_pair = first_pair, pf_chunk = lo - first_pair * nchunk, pf_I = 0, pf_J = 1;
    if (nitems > 0) rr_pair(nb, step, pf_pair, pf_I, pf_J);
    auto load_next = [&](int st) {
        c128* ps = Ps + (size_t)st * JP * G_PITCH;
        const c128* src = Z + (int64_t)pf_chunk * G_BKR + l_row;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int col = l_col0 + 8 * i;
            cp_async16(ps + col * G_PITCH + l_row, src + panel_col(pf_I, pf_J, col) * ldz, true);
        }
        if (++pf_chunk == nchunk) {
            pf_chunk = 0;
            ++pf_pair;
            if (pf_pair < npairs) rr_pair(nb, step, pf_pair, pf_I, pf_J);
        }
    };

    double cr[5][2], ci[5][2], cs[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int i = 0; i < 5; ++i) cr[i][0] = cr[i][1] = ci[i][0] = ci[i][1] = 0.0;

#pragma unroll
    for (int s = 0; s < G_NST - 1; ++s) {
        if (s < nitems) load_next(s);
        cp_async_commit();
    }
    int pair = first_pair, chunk = lo - first_pair * nchunk;
    for (int it = 0; it < nitems; ++it) {
        cp_async_wait<G_NST - 2>();
        __syncthreads();
        {
            int nx = it + G_NST - 1;
            if (nx < nitems) load_next(nx % G_NST);
            cp_async_commit();
        }
        const c128* ps = Ps + (size_t)(it % G_NST) * JP * G_PITCH;
        switch (warp) {
            case 0: gram_mma_chunk<0, CROSS>(ps, g, t, cr, ci, cs); break;
            case 1: gram_mma_chunk<1, CROSS>(ps, g, t, cr, ci, cs); break;
            case 2: gram_mma_chunk<2, CROSS>(ps, g, t, cr, ci, cs); break;
            case 3: gram_mma_chunk<3, CROSS>(ps, g, t, cr, ci, cs); break;
            case 4: gram_mma_chunk<4, CROSS>(ps, g, t, cr, ci, cs); break;
            case 5: gram_mma_chunk<5, CROSS>(ps, g, t, cr, ci, cs); break;
            case 6: gram_mma_chunk<6, CROSS>(ps, g, t, cr, ci, cs); break;
            default: gram_mma_chunk<7, CROSS>(ps, g, t, cr, ci, cs); break;
        }
        const bool last_of_pair = (it + 1 == nitems) || (chunk + 1 == nchunk);
        if (last_of_pair) {
            c128* out = Gpart + ((size_t)2 * blockIdx.x + (pair != first_pair ? 1 : 0)) * (JP * JP);
            switch (warp) {
                case 0: gram_flush<0, CROSS>(out, g, t, cr, ci, cs); break;
                case 1: gram_flush<1, CROSS>(out, g, t, cr, ci, cs); break;
                case 2: gram_flush<2, CROSS>(out, g, t, cr, ci, cs); break;
                case 3: gram_flush<3, CROSS>(out, g, t, cr, ci, cs); break;
                case 4: gram_flush<4, CROSS>(out, g, t, cr, ci, cs); break;
                case 5: gram_flush<5, CROSS>(out, g, t, cr, ci, cs); break;
                case 6: gram_flush<6, CROSS>(out, g, t, cr, ci, cs); break;
                default: gram_flush<7, CROSS>(out, g, t, cr, ci, cs); break;
            }
        }
        if (++chunk == nchunk) {
            chunk = 0;
            ++pair;
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// low-precision cross Gram (mixed-precision Jacobi): G_IJ = X_I^H X_J from the FP32 shadow copy X32 of X on the
// TF32 tensor path (mma.sync m16n8k8, FP32 accumulate).  The rotations W_p that the evd kernel derives from a Gram
// matrix are exact plane rotations whatever the accuracy of that Gram matrix, and X <- X W_p is applied in FP64, so
// the singular values are untouched by the Gram precision; only the QUALITY of the rotations depends on it.  While
// the couplings |g_ij| / (|x_i||x_j|) are still >~ 1e-2 (the linear phase: 7-8 of ~11 sweeps) a Gram matrix with
// ~1e-4 relative noise steers the rotations just as well, at half the bytes and none of the FP64 tensor pipe.
// The last sweeps (quadratic phase, convergence test) always use the FP64 Gram kernel above.
// Same work partition, partial-slot convention and cross-block flush as jacobi_gram_kernel<1>.
constexpr int G32_PITCH = G_BKR + 4;  // in float2: 4 k x 4 m fragment loads of a half warp hit 16 distinct 8-byte banks
constexpr size_t GRAM32_SMEM = (size_t)G_NST * JP * G32_PITCH * sizeof(float2);

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void hmma1688_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(256, 2)
    jacobi_gram32_kernel(const float2* __restrict__ Z32, int64_t ldz, int mp, int nb, int step, int npairs,
                         c128* __restrict__ Gpart) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Ps = reinterpret_cast<float2*>(smem_raw);  // [stage][panel column][row], pitch G32_PITCH
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int mi = warp >> 2, nj = warp & 3;  // warp tile: rows (columns of block I) mi*16.., columns (of block J) nj*8..
    const int nchunk = mp / G_BKR, total = npairs * nchunk;
    int lo, hi;
    gram_item_range(blockIdx.x, gridDim.x, total, lo, hi);
    const int nitems = hi - lo;

    // loader: 16-byte cp.async = two consecutive rows of one column; 64 columns x 16 row pairs = 1024 / 256 threads
    const int l_rp = tid & 15, l_col0 = tid >> 4;
    const int first_pair = (nitems > 0) ? lo / nchunk : 0;
    int pf_pair = first_pair, pf_chunk = lo - first_pair * nchunk, pf_I = 0, pf_J = 1;
    if (nitems > 0) rr_pair(nb, step, pf_pair, pf_I, pf_J);
    auto load_next = [&](int st) {
        float2* ps = Ps + (size_t)st * JP * G32_PITCH;
        const float2* src = Z32 + (int64_t)pf_chunk * G_BKR + 2 * l_rp;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int col = l_col0 + 16 * i;
            cp_async16(ps + col * G32_PITCH + 2 * l_rp, src + panel_col(pf_I, pf_J, col) * ldz, true);
        }
        if (++pf_chunk == nchunk) {
            pf_chunk = 0;
            ++pf_pair;
            if (pf_pair < npairs) rr_pair(nb, step, pf_pair, pf_I, pf_J);
        }
    };

    float cr[4] = {0.f, 0.f, 0.f, 0.f}, ci[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < G_NST - 1; ++s) {
        if (s < nitems) load_next(s);
        cp_async_commit();
    }
    int pair = first_pair, chunk = lo - first_pair * nchunk;
    for (int it = 0; it < nitems; ++it) {
        cp_async_wait<G_NST - 2>();
        __syncthreads();
        {
            int nx = it + G_NST - 1;
            if (nx < nitems) load_next(nx % G_NST);
            cp_async_commit();
        }
        const float2* pa = Ps + (size_t)(it % G_NST) * JP * G32_PITCH + (mi * 16 + g) * G32_PITCH + t;
        const float2* pb = Ps + (size_t)(it % G_NST) * JP * G32_PITCH + (JB + nj * 8 + g) * G32_PITCH + t;
#pragma unroll
        for (int kk = 0; kk < G_BKR / 8; ++kk) {
            uint32_t ar[4], ai[4], nai[4], br[2], bi[2];
#pragma unroll
            for (int e = 0; e < 4; ++e) {  // a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]; A = X_I^H
                float2 v = pa[(e & 1) * 8 * G32_PITCH + kk * 8 + 4 * (e >> 1)];
                ar[e] = to_tf32(v.x);
                ai[e] = to_tf32(v.y);
                nai[e] = ai[e] ^ 0x80000000u;
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {  // b0 = B[t][g], b1 = B[t+4][g]; B = X_J
                float2 v = pb[kk * 8 + 4 * e];
                br[e] = to_tf32(v.x);
                bi[e] = to_tf32(v.y);
            }
            // conj(a) b = (ar br + ai bi) + i (ar bi - ai br)
            hmma1688_tf32(cr, ar, br);
            hmma1688_tf32(cr, ai, bi);
            hmma1688_tf32(ci, ar, bi);
            hmma1688_tf32(ci, nai, br);
        }
        const bool last_of_pair = (it + 1 == nitems) || (chunk + 1 == nchunk);
        if (last_of_pair) {
            c128* out = Gpart + ((size_t)2 * blockIdx.x + (pair != first_pair ? 1 : 0)) * (JP * JP);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const int row = mi * 16 + g + 8 * (h >> 1), col = JB + nj * 8 + 2 * t + (h & 1);
                out[row + JP * col] = make_double2((double)cr[h], (double)ci[h]);
                out[col + JP * row] = make_double2((double)cr[h], -(double)ci[h]);
                cr[h] = 0.f;
                ci[h] = 0.f;
            }
        }
        if (++chunk == nchunk) {
            chunk = 0;
            ++pair;
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Mixed-precision Jacobi, low-precision stage ("stage A", see qb_svd_factor): the iteration runs on an FP32 shadow
// S = [X32; V32] (2 mp x np, column-major float2, lds = 2 mp: the image of X on top, the accumulated rotations below).
// Rotations still come from jacobi_evd_kernel (exact plane rotations in FP64 from whatever Gram it is given).

// S <- [narrow(X); I]
__global__ void lp_init_kernel(const c128* __restrict__ Z, int64_t ldz, int mp, int np, float2* __restrict__ S,
                               int64_t lds) {
    const int64_t total = (int64_t)2 * mp * np;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx % (2 * mp), c = idx / (2 * mp);
        float2 v;
        if (r < mp) {
            const c128 z = Z[r + c * ldz];
            v = make_float2((float)z.x, (float)z.y);
        } else {
            v = make_float2((r - mp) == c ? 1.f : 0.f, 0.f);
        }
        S[r + c * lds] = v;
    }
}

// V (np x np, ld = np, ComplexF64) <- lower half of S
__global__ void lp_extract_v_kernel(const float2* __restrict__ S, int64_t lds, int mp, int np, c128* __restrict__ V) {
    const int64_t total = (int64_t)np * np;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = idx % np, c = idx / np;
        const float2 v = S[mp + r + c * lds];
        V[idx] = make_double2((double)v.x, (double)v.y);
    }
}

__global__ void add_diag_kernel(c128* __restrict__ T, int n, double v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) T[(int64_t)i * n + i].x += v;
}

// Full 64 x 64 Gram of every pair of a step (step 0 of a sweep) from X32, FP32 FMA.  grid (nsplit, npairs):
// CTA (s, p) sums the rows [s rows_per, (s + 1) rows_per) of the panel of pair p and writes partial slot
// 2 (p nsplit + s) -- what jacobi_evd_kernel expects from gram_ctas = npairs nsplit CTAs with nchunk = nsplit.
// Thread (ti, tj) of the 16 x 16 grid owns the 4 x 4 block G[4 ti .. , 4 tj ..].
__global__ void __launch_bounds__(256) lp_gram_full_kernel(const float2* __restrict__ S, int64_t lds, int rows_per, int nb,
                                                           int step, c128* __restrict__ Gpart) {
    __shared__ float2 Xs[32][JP + 1];
    const int tid = threadIdx.x, ti = tid & 15, tj = tid >> 4;
    const int pair = blockIdx.y, split = blockIdx.x, nsplit = gridDim.x;
    int I, J;
    rr_pair(nb, step, pair, I, J);
    const float2* src = S + (int64_t)split * rows_per;
    float gr[4][4], gi[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) gr[a][b] = gi[a][b] = 0.f;
    for (int r0 = 0; r0 < rows_per; r0 += 32) {
        __syncthreads();
        for (int e = tid; e < 32 * JP; e += 256) {
            const int r = e & 31, c = e >> 5;
            Xs[r][c] = src[r0 + r + panel_col(I, J, c) * lds];
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            float2 a[4], b[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                a[x] = Xs[r][4 * ti + x];
                b[x] = Xs[r][4 * tj + x];
            }
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) {  // conj(a) b
                    gr[x][y] += a[x].x * b[y].x + a[x].y * b[y].y;
                    gi[x][y] += a[x].x * b[y].y - a[x].y * b[y].x;
                }
        }
    }
    c128* out = Gpart + (size_t)2 * (pair * nsplit + split) * (JP * JP);
#pragma unroll
    for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) out[(4 * ti + x) + JP * (4 * tj + y)] = make_double2((double)gr[x][y], (double)gi[x][y]);
}

// max |a - b| and max |b| (as float bit patterns: non-negative floats order like unsigned integers)
__global__ void lp_maxdiff_kernel(const float2* __restrict__ a, const float2* __restrict__ b, int64_t n, unsigned* __restrict__ out) {
    float d = 0.f, m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        d = fmaxf(d, fmaxf(fabsf(a[i].x - b[i].x), fabsf(a[i].y - b[i].y)));
        m = fmaxf(m, fmaxf(fabsf(b[i].x), fabsf(b[i].y)));
    }
    atomicMax(out, __float_as_uint(d));
    atomicMax(out + 1, __float_as_uint(m));
}

// Reference update of the shadow, FP32 FMA: S_p <- S_p W_p on all 2 mp rows.  grid (rows / 32, npairs).  Slow (CUDA
// cores); kept as the specification the tcgen05 kernel is tested against (QB200_LP_UPDATE=simple).
__global__ void __launch_bounds__(256) lp_update_simple_kernel(float2* __restrict__ S, int64_t lds, int nb, int step,
                                                               const c128* __restrict__ Wg, const int* __restrict__ flags) {
    __shared__ float2 Ws[JP * JP];   // [k][n]
    __shared__ float2 Xs[JP * 32];   // [k][row]
    const int pair = blockIdx.y;
    if (!flags[pair]) return;
    int I, J;
    rr_pair(nb, step, pair, I, J);
    const int tid = threadIdx.x, r = tid & 31, cg = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * 32;
    for (int e = tid; e < JP * JP; e += 256) {
        const int k = e & 63, n = e >> 6;
        const c128 w = Wg[(size_t)pair * (JP * JP) + n * JP + k];
        Ws[k * JP + n] = make_float2((float)w.x, (float)w.y);
    }
    for (int e = tid; e < JP * 32; e += 256) {
        const int rr = e & 31, k = e >> 5;
        Xs[k * 32 + rr] = S[row0 + rr + panel_col(I, J, k) * lds];
    }
    __syncthreads();
    float2 acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = make_float2(0.f, 0.f);
    for (int k = 0; k < JP; ++k) {
        const float2 x = Xs[k * 32 + r];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 w = Ws[k * JP + cg * 8 + j];
            acc[j].x += x.x * w.x - x.y * w.y;
            acc[j].y += x.x * w.y + x.y * w.x;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) S[row0 + r + panel_col(I, J, cg * 8 + j) * lds] = acc[j];
}

// ---------------------------------------------------------------------------------------------
// evd kernel: two-sided Jacobi on the 64 x 64 Hermitian Gram matrix of one block pair.
//
// Each parallel step applies 32 disjoint plane rotations J = prod_k J_k:
//   phase 1 (32 threads): rotation k from (g_pp, g_qq, g_pq), rsqrt-based (no fp64 divide/sqrt chains);
//   phase 2 (all threads): W <- W J (2048 column-pair items) and G <- J^H G J done per 2 x 2 block
//            B_kl = G[{p_k,q_k},{p_l,q_l}] -> J_k^H B_kl J_l, upper blocks only (k <= l), mirrored by Hermitian
//            symmetry: every block touches only its own entries, so the update is in place with ONE barrier.
// Orderings: mode 0 = round robin over `nact` columns (a lone pair = the whole small matrix, run to
// convergence); mode 1 = within-block pairs (31 steps) then cross pairs (32 steps); mode 2 = cross pairs only
// (i, 32 + (i + t) mod 32): the blocks of a pair were made orthogonal internally at step 0 of the sweep, so
// the later steps of a sweep only need the cross rotations -- one outer sweep then rotates every column pair
// exactly once, like scalar cyclic Jacobi, at a quarter of the shared-memory work of a full inner solve.
// 17 warps: warp 0 prepares rotations while 16 warps = 512 threads apply them -- exactly 4 W items per thread -- and
// the 528 upper 2 x 2 blocks of the G update fit in ONE pass (with 512 threads 16 of them did two: the step took
// two rounds of a latency-bound block update).
constexpr int EVD_THREADS = 544;
constexpr int EVD_LOADERS = 512;  // threads that assemble G (8 elements each)
constexpr int EVD_NBLK = 32 * 33 / 2;
constexpr size_t EVD_SMEM_TAIL = 32 * sizeof(double) + 32 * sizeof(c128) + 64 * sizeof(int) + 64 * sizeof(double) +
                                 EVD_NBLK * sizeof(short);
constexpr size_t EVD_SMEM = (size_t)2 * JP * GLD * sizeof(c128) + EVD_SMEM_TAIL;
constexpr size_t EVD_SMEM_WREG = (size_t)JP * GLD * sizeof(c128) + EVD_SMEM_TAIL;  // W in registers: G only

// WREG = 1 (mode 2 only: 32 cross rotations (i, 32 + (i + t) mod 32) per step, one inner sweep): W never touches shared
// memory.  Warp w of the 16 apply-warps owns rows 4w .. 4w+3; lane i keeps, for each of them, W[r][i] and the entry of
// the column currently paired with i, W[r][32 + (i + t) mod 32]: the rotation of lane i acts on exactly these two, and
// the partner column of the next step is the one lane i+1 holds now -- one warp shuffle per step passes it on.  After
// the 32 steps every lane is back at column 32 + i.  Shared memory drops to the 65 KB of G (the kernel then shares an SM
// with an update / Gram CTA of another stream) and the W phase to register arithmetic.
template <int WREG>
__global__ void __launch_bounds__(EVD_THREADS, 1)
    jacobi_evd_kernel(const c128* __restrict__ Gpart, int gram_ctas, int nchunk, int npairs_total,
                      c128* __restrict__ Wout, int* __restrict__ flags,
                      unsigned long long* __restrict__ sweep_stat, double rot_tol, int inner_sweeps,
                      const double* __restrict__ scale_in, unsigned long long* __restrict__ scale_out, double abs_c,
                      int nact, int mode, c128* __restrict__ Dstore, int nb, int step) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* G = reinterpret_cast<c128*>(smem_raw);
    c128* W = G + JP * GLD;                                   // WREG: not allocated, never touched
    double* rcs = reinterpret_cast<double*>(WREG ? W : W + JP * GLD);   // rotation cosines, buffer 0
    double* red = rcs + 32 + 64 + 32;                          // reduction scratch (64 doubles), later cosines buffer 1
    short* tri = reinterpret_cast<short*>(red + 64);

    const int tid = threadIdx.x, pair = blockIdx.x;
    // partial Grams of this pair: the gram CTAs whose item range overlaps [pair*nchunk, (pair+1)*nchunk), in order
    __shared__ int slots[64];
    __shared__ int nslots;
    if (tid == 0) {
        const int total_items = npairs_total * nchunk;
        const int p_lo = pair * nchunk, p_hi = p_lo + nchunk;
        int c = (int)(((long long)p_lo * gram_ctas) / total_items) - 1;
        if (c < 0) c = 0;
        int ns = 0;
        for (; c < gram_ctas && ns < 64; ++c) {
            int lo, hi;
            gram_item_range(c, gram_ctas, total_items, lo, hi);
            if (lo >= p_hi) break;
            if (hi <= p_lo || hi == lo) continue;
            slots[ns++] = 2 * c + ((lo / nchunk) != pair ? 1 : 0);
        }
        nslots = ns;
    }
    __syncthreads();
    // mode 2 (cross rotations only): the gram kernel computed just the I x J cross block; the two 32 x 32 diagonal
    // blocks are the ones this kernel left in Dstore when these column blocks were last rotated (G <- W^H G W is
    // carried along exactly by the two-sided update; refreshed from a full Gram at step 0 of every sweep).
    int blkI = 0, blkJ = 0;
    if (Dstore) rr_pair(nb, step, pair, blkI, blkJ);
    if (tid < EVD_LOADERS) {
        // 8 elements per thread; the loads of one slot are issued together (memory-level parallelism), slots in order
        constexpr int PER = JP * JP / EVD_LOADERS;
        double sx[PER], sy[PER];
        bool from_d[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = tid + EVD_LOADERS * j, er = e & 63, ec = e >> 6;
            from_d[j] = (mode == 2) && ((er < JB) == (ec < JB));
            sx[j] = sy[j] = 0.0;
            if (from_d[j]) {
                c128 v = Dstore[(size_t)(er < JB ? blkI : blkJ) * (JB * JB) + (er & 31) + JB * (ec & 31)];
                sx[j] = v.x;
                sy[j] = v.y;
            }
        }
        for (int i = 0; i < nslots; ++i) {
            const c128* src = Gpart + (size_t)slots[i] * (JP * JP);
            c128 v[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j)
                v[j] = from_d[j] ? make_double2(0.0, 0.0) : src[tid + EVD_LOADERS * j];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                sx[j] += v[j].x;
                sy[j] += v[j].y;
            }
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = tid + EVD_LOADERS * j, r = e & 63, c = e >> 6;
            G[c * GLD + r] = make_double2(sx[j], sy[j]);
            if constexpr (!WREG) W[c * GLD + r] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        }
    }
    for (int e = tid; e < EVD_NBLK; e += EVD_THREADS) {
        // e -> (k, l), k <= l, row-major over the upper triangle of a 32 x 32 grid
        int k = 0, rem = e;
        while (rem >= 32 - k) {
            rem -= 32 - k;
            ++k;
        }
        tri[e] = (short)((k << 8) | (k + rem));
    }
    __syncthreads();
    // A pair (i, j) counts as orthogonal when |g_ij| <= rot_tol |x_i||x_j| or |g_ij| <= abs_tol max(|x_i|, |x_j|):
    // the first is the usual relative criterion, the second the accuracy a GEMM-applied rotation can deliver at
    // all (every column of an updated panel carries an absolute error ~ eps * sigma_1); without it columns whose
    // norm is at the rounding level of sigma_1 would be rotated for ever.
    // abs_c < 0: the factor lives in scale_in[2] (set by the host between sweeps without re-capturing the sweep graph)
    const double abs_tol = (abs_c < 0.0 ? scale_in[2] : abs_c) * scale_in[0];
    const double rt2 = rot_tol * rot_tol, at2 = abs_tol * abs_tol;
    double mx = 0.0, gmax = 0.0;
    for (int e = tid; e < JP * JP; e += EVD_THREADS) {
        int r = e & 63, c = e >> 6;
        if (r == c) gmax = fmax(gmax, G[c * GLD + c].x);
        if (r < c) {
            double a = G[r * GLD + r].x, b = G[c * GLD + c].x;
            c128 v = G[c * GLD + r];
            double d = a * b, n2 = v.x * v.x + v.y * v.y;
            if (d > 0.0 && n2 > fmax(rt2 * d, at2 * fmax(a, b))) mx = fmax(mx, n2 / d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) {
        if (gmax > 0.0) atomicMax(scale_out, (unsigned long long)__double_as_longlong(sqrt(gmax)));
        red[tid >> 5] = mx;
    }
    __syncthreads();
    if (tid < 32) {
        double v = (tid < EVD_THREADS / 32) ? red[tid] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (tid == 0) {
            v = sqrt(v);
            red[32] = v;
            atomicMax(sweep_stat, (unsigned long long)__double_as_longlong(v));
        }
    }
    __syncthreads();
    mx = red[32];
    if (!(mx > 0.0)) {
        if (tid == 0) flags[pair] = 0;
        if (Dstore && mode == 1)  // fresh diagonal blocks for the cross-only steps that follow
            for (int e = tid; e < 2 * JB * JB; e += EVD_THREADS) {
                int h = e >> 10, r = e & 31, c = (e >> 5) & 31;
                Dstore[(size_t)(h ? blkJ : blkI) * (JB * JB) + r + JB * c] = G[(c + JB * h) * GLD + r + JB * h];
            }
        return;
    }
    if (tid == 0) flags[pair] = 1;

    // Software pipeline over the flattened steps g = (inner sweep, step): the rotations of step g+1 only need the
    // UPDATED G, not W, so once G <- J^H G J is done (all threads) warp 0 computes the next rotations -- a long
    // dependent FP64 chain -- while the other 15 warps apply the current rotations to W.
    const int nsteps_rr = (nact == 2) ? 1 : nact - 1;
    const int nst = (mode == 0) ? nsteps_rr : (mode == 1 ? 63 : 32);
    const int nrot = (mode == 0) ? nact / 2 : 32;
    const int total_steps = inner_sweeps * nst;
    __shared__ int rot_in_sweep[16];
    if (tid < 16) rot_in_sweep[tid] = 0;
    // rotation buffers: [2][32]
    double* rcs2 = rcs;             // reuse: rcs[0..31] buffer 0, red[0..31] is free after the reduction -> buffer 1
    double* rcsb[2] = {rcs2, red};
    __shared__ c128 rsn_buf[2][32];
    __shared__ int rpq_buf[2][64];
    auto compute_rotations = [&](int gstep, int buf) {
        // executed by warp 0 only (tid < 32)
        const int st = gstep % nst;
        int p = 0, q = 0;
        bool active = true;
        if (mode == 0) {
            active = tid < nact / 2;
            if (active) rr_pair(nact, st, tid, p, q);
        } else if (mode == 2 || st >= 31) {
            int t = (mode == 2) ? st : st - 31;
            p = tid;
            q = 32 + ((tid + t) & 31);
        } else {
            int half = tid >> 4;
            rr_pair(32, st, tid & 15, p, q);
            p += 32 * half;
            q += 32 * half;
        }
        double cs = 1.0;
        c128 sn = make_double2(0.0, 0.0);
        if (active) {
            double a = G[p * GLD + p].x, b = G[q * GLD + q].x;
            c128 c = G[q * GLD + p];  // G[p][q] = x_p^H x_q  (row p, column q)
            double n2 = c.x * c.x + c.y * c.y;
            if (n2 > fmax(rt2 * a * b, at2 * fmax(a, b)) && n2 > 0.0) {
                double inv = rsqrt(n2);
                double zeta = 0.5 * (b - a) * inv;
                double r2 = 1.0 + zeta * zeta;
                double rr = r2 * rsqrt(r2);
                double tt = copysign(1.0, zeta) / (fabs(zeta) + rr);
                cs = rsqrt(1.0 + tt * tt);
                double sv = tt * cs * inv;
                sn = make_double2(sv * c.x, sv * c.y);  // sin * c/|c|
                rot_in_sweep[gstep / nst] = 1;
            }
        }
        rcsb[buf][tid] = cs;
        rsn_buf[buf][tid] = sn;
        rpq_buf[buf][2 * tid] = p;
        rpq_buf[buf][2 * tid + 1] = q;
    };
    c128 wp[4], wq[4];  // WREG: see the kernel comment
    if constexpr (WREG) {
        const int lane = tid & 31, r0 = 4 * ((tid >> 5) - 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            wp[j] = make_double2(r0 + j == lane ? 1.0 : 0.0, 0.0);
            wq[j] = make_double2(r0 + j == 32 + lane ? 1.0 : 0.0, 0.0);
        }
    }
    __syncthreads();
    if (tid < 32) compute_rotations(0, 0);
    __syncthreads();
    for (int gs = 0; gs < total_steps; ++gs) {
        const int buf = gs & 1;
        if (gs > 0 && gs % nst == 0 && !rot_in_sweep[gs / nst - 1]) break;  // previous inner sweep was all identity
        const double* rc = rcsb[buf];
        const c128* rs = rsn_buf[buf];
        const int* rp = rpq_buf[buf];
        // phase A: G <- J^H G J per 2 x 2 block (k <= l), mirrored
        for (int e = tid; e < EVD_NBLK; e += EVD_THREADS) {
            int k = tri[e] >> 8, l = tri[e] & 0xff;
            if (l >= nrot) continue;
            c128 sk = rs[k], sl = rs[l];
            bool rk = (sk.x != 0.0 || sk.y != 0.0), rl = (sl.x != 0.0 || sl.y != 0.0);
            if (!rk && !rl) continue;
            double ck = rc[k], cl = rc[l];
            int pk = rp[2 * k], qk = rp[2 * k + 1], pl = rp[2 * l], ql = rp[2 * l + 1];
            c128 b00 = G[pl * GLD + pk], b01 = G[ql * GLD + pk], b10 = G[pl * GLD + qk], b11 = G[ql * GLD + qk];
            c128 t00 = csub(cscale(b00, cl), cmul(cconj(sl), b01));
            c128 t01 = cadd(cmul(sl, b00), cscale(b01, cl));
            c128 t10 = csub(cscale(b10, cl), cmul(cconj(sl), b11));
            c128 t11 = cadd(cmul(sl, b10), cscale(b11, cl));
            b00 = csub(cscale(t00, ck), cmul(sk, t10));
            b10 = cadd(cmul(cconj(sk), t00), cscale(t10, ck));
            b01 = csub(cscale(t01, ck), cmul(sk, t11));
            b11 = cadd(cmul(cconj(sk), t01), cscale(t11, ck));
            if (k == l) {
                b00.y = 0.0;
                b11.y = 0.0;
                b01 = make_double2(0.0, 0.0);  // annihilated by construction
                b10 = make_double2(0.0, 0.0);
                G[pk * GLD + pk] = b00;
                G[qk * GLD + qk] = b11;
                G[qk * GLD + pk] = b01;
                G[pk * GLD + qk] = b10;
            } else {
                G[pl * GLD + pk] = b00;
                G[ql * GLD + pk] = b01;
                G[pl * GLD + qk] = b10;
                G[ql * GLD + qk] = b11;
                G[pk * GLD + pl] = cconj(b00);
                G[pk * GLD + ql] = cconj(b01);
                G[qk * GLD + pl] = cconj(b10);
                G[qk * GLD + ql] = cconj(b11);
            }
        }
        __syncthreads();
        // phase B: warp 0 prepares the next step's rotations, warps 1..15 apply the current ones to W
        if (tid < 32) {
            if (gs + 1 < total_steps) compute_rotations(gs + 1, buf ^ 1);
        } else if constexpr (WREG) {
            const int lane = tid & 31;
            const c128 sn = rs[lane];
            const double cs = rc[lane];
            if (sn.x != 0.0 || sn.y != 0.0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const c128 xp = wp[j], xq = wq[j];
                    wp[j] = csub(cscale(xp, cs), cmul(cconj(sn), xq));
                    wq[j] = cadd(cmul(sn, xp), cscale(xq, cs));
                }
            }
            const int from = (lane + 1) & 31;  // the column that pairs with this lane's p column at the next step
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                wq[j].x = __shfl_sync(0xffffffffu, wq[j].x, from);
                wq[j].y = __shfl_sync(0xffffffffu, wq[j].y, from);
            }
        } else {
            for (int item = tid - 32; item < 32 * JP; item += EVD_THREADS - 32) {
                int r = item & 63, k = item >> 6;
                if (k >= nrot) continue;
                c128 sn = rs[k];
                if (sn.x == 0.0 && sn.y == 0.0) continue;
                double cs = rc[k];
                int p = rp[2 * k], q = rp[2 * k + 1];
                c128 xp = W[p * GLD + r], xq = W[q * GLD + r];
                W[p * GLD + r] = csub(cscale(xp, cs), cmul(cconj(sn), xq));
                W[q * GLD + r] = cadd(cmul(sn, xp), cscale(xq, cs));
            }
        }
        __syncthreads();
    }
    c128* dst = Wout + (size_t)pair * (JP * JP);
    if constexpr (WREG) {
        if (tid >= 32) {  // after the 32 steps lane i holds columns i and 32 + i of its 4 rows
            const int lane = tid & 31, r0 = 4 * ((tid >> 5) - 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                dst[lane * JP + r0 + j] = wp[j];
                dst[(32 + lane) * JP + r0 + j] = wq[j];
            }
        }
    } else {
        for (int e = tid; e < JP * JP; e += EVD_THREADS) {
            int r = e & 63, c = e >> 6;
            dst[e] = W[c * GLD + r];
        }
    }
    if (Dstore)
        for (int e = tid; e < 2 * JB * JB; e += EVD_THREADS) {
            int h = e >> 10, r = e & 31, c = (e >> 5) & 31;
            Dstore[(size_t)(h ? blkJ : blkI) * (JB * JB) + r + JB * c] = G[(c + JB * h) * GLD + r + JB * h];
        }
}

// ---------------------------------------------------------------------------------------------
// update kernel: X_p <- X_p W_p in place.  Work items are (pair, 32-row chunk), flattened pair-major and cut into
// equal contiguous ranges over a persistent grid of 2 x #SM CTAs (2048 items over 296 CTAs at chi = 1024: no tail
// wave).  Each of the 8 warps owns 8 output columns and keeps ITS 64 x 8 slice of W_p in registers as DMMA B
// fragments (16 complex per lane) for as long as the CTA stays on the pair, so shared memory holds nothing but
// the 3-stage cp.async ring of chunks (102 KB): two CTAs fit on an SM and the prologue, barrier and store phases of
// one overlap the DMMA stream of the other.  Results go straight from the accumulators to global memory.
constexpr int U_ROWS = 32, U_ZP = 34, U_NST = 3;  // pitch = 2 mod 8 (in c128): conflict-free 16-byte fragment loads
constexpr size_t UPD_SMEM = (size_t)U_NST * JP * U_ZP * sizeof(c128);
constexpr int UPD_THREADS = 256;

// VAR != 0: timing variants for qb_update_bench only (wrong results): bit 0 no operand-sum DADDs, bit 1 no global
// stores, bit 2 no cp.async after the prologue, bit 3 no barrier
template <int M3, int VAR = 0>
__global__ void __launch_bounds__(UPD_THREADS, 2)
    jacobi_update_kernel(c128* __restrict__ Z, int64_t ldz, int nb, int step, const c128* __restrict__ Wg,
                         const int* __restrict__ flags, int npairs, int nchunk, float2* __restrict__ Z32) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* Zs = reinterpret_cast<c128*>(smem_raw);  // [stage][col][row] pitch 34
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int total = npairs * nchunk;
    int lo, hi;
    gram_item_range(blockIdx.x, gridDim.x, total, lo, hi);

    // The items of a CTA are walked by two cursors (compute, and prefetch U_NST - 1 items ahead) that carry
    // (pair, chunk, I, J) incrementally: no division and no global load on the critical path.  The "pair is already
    // orthogonal (W = I)" flags of the <= 32 pairs a range can touch are fetched once into a bit mask.
    const int p0 = (lo < hi) ? lo / nchunk : 0;
    unsigned amask;
    {
        const int p = p0 + lane;
        const int f = (lo < hi && p < npairs) ? flags[p] : 0;
        amask = __ballot_sync(0xffffffffu, f != 0);
    }
    struct Cursor {
        int item, pair, chunk, I, J;
    };
    auto pair_active = [&](int pair) {
        const int d = pair - p0;
        return d < 32 ? ((amask >> d) & 1u) != 0 : flags[pair] != 0;
    };
    auto settle = [&](Cursor& c) {  // skip the pairs with W = I, then resolve the block indices of the pair
        while (c.item < hi && !pair_active(c.pair)) {
            c.item += nchunk - c.chunk;
            c.chunk = 0;
            ++c.pair;
        }
        if (c.item < hi) rr_pair(nb, step, c.pair, c.I, c.J);
    };
    auto advance = [&](Cursor& c) {
        ++c.item;
        if (++c.chunk == nchunk) {
            c.chunk = 0;
            ++c.pair;
            settle(c);
        }
    };
    const int l_row = tid & 31, l_col0 = tid >> 5;
    auto load_chunk = [&](const Cursor& c, int st) {
        c128* zs = Zs + (size_t)st * JP * U_ZP;
        const c128* src = Z + (int64_t)c.chunk * U_ROWS + l_row;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int col = l_col0 + 8 * i;
            cp_async16(zs + col * U_ZP + l_row, src + panel_col(c.I, c.J, col) * ldz, true);
        }
    };

    Cursor cur{lo, p0, lo - p0 * nchunk, 0, 1};
    settle(cur);
    Cursor pf = cur;
#pragma unroll
    for (int s = 0; s < U_NST - 1; ++s) {
        if (pf.item < hi) {
            load_chunk(pf, s);
            advance(pf);
        }
        cp_async_commit();
    }
    int cur_pair = -1, stage = 0;
    c128 breg[JP / 4];  // W[kk*4 + t][warp*8 + g], kk = 0..15
    while (cur.item < hi) {
        if (cur.pair != cur_pair) {
            const c128* wsrc = Wg + (size_t)cur.pair * (JP * JP) + (warp * 8 + g) * JP + t;
#pragma unroll
            for (int kk = 0; kk < JP / 4; ++kk) breg[kk] = wsrc[kk * 4];
            cur_pair = cur.pair;
        }
        // ONE barrier per chunk: it publishes the landed chunk AND says every warp has left the previous chunk,
        // whose buffer the prefetch below overwrites
        cp_async_wait<U_NST - 2>();
        if (!(VAR & 8)) __syncthreads();
        if (pf.item < hi) {
            if (!(VAR & 4)) load_chunk(pf, (stage + U_NST - 1) % U_NST);
            advance(pf);
        }
        cp_async_commit();
        const c128* za = Zs + (size_t)stage * JP * U_ZP + g;
        const int64_t r0 = (int64_t)cur.chunk * U_ROWS + g;
        if constexpr (M3) {
            // 3M complex product (Gauss): P = Ar Br, Q = Ai Bi, S = (Ar + Ai)(Br + Bi); Cr = P - Q, Ci = S - P - Q.
            // 3 DMMA + the operand sums instead of 4 DMMA per complex 8x8x4 tile; the error stays norm-wise
            // eps (|Ar| + |Ai|)(|Br| + |Bi|), the class the Jacobi update needs.  Two 16-row halves keep the three
            // accumulator sets inside the 128-register budget of 2 CTAs/SM.
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                double pp[2][2], qq[2][2], ss[2][2];
#pragma unroll
                for (int a = 0; a < 2; ++a) pp[a][0] = pp[a][1] = qq[a][0] = qq[a][1] = ss[a][0] = ss[a][1] = 0.0;
#pragma unroll
                for (int kk = 0; kk < JP / 4; ++kk) {
                    const double br = breg[kk].x, bi = breg[kk].y, bs = (VAR & 1) ? br : br + bi;
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        c128 v = za[(kk * 4 + t) * U_ZP + (hh * 2 + a) * 8];
                        dmma884(pp[a], v.x, br);
                        dmma884(qq[a], v.y, bi);
                        dmma884(ss[a], (VAR & 1) ? v.x : v.x + v.y, bs);
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int64_t zo = r0 + panel_col(cur.I, cur.J, warp * 8 + 2 * t + h) * ldz;
#pragma unroll
                    for (int a = 0; a < 2; ++a) {
                        const double re = pp[a][h] - qq[a][h], im = ss[a][h] - pp[a][h] - qq[a][h];
                        if ((VAR & 2) && re != 1.2345e300) continue;
                        Z[zo + (hh * 2 + a) * 8] = make_double2(re, im);
                        if (Z32) Z32[zo + (hh * 2 + a) * 8] = make_float2((float)re, (float)im);  // FP32 shadow (gram32)
                    }
                }
            }
        } else {
            double cr[4][2], ci[4][2];
#pragma unroll
            for (int a = 0; a < 4; ++a) cr[a][0] = cr[a][1] = ci[a][0] = ci[a][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < JP / 4; ++kk) {
                const double br = breg[kk].x, bi = breg[kk].y;
                double ar[4], ai[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    c128 v = za[(kk * 4 + t) * U_ZP + a * 8];
                    ar[a] = v.x;
                    ai[a] = v.y;
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    dmma884(cr[a], ar[a], br);
                    dmma884(ci[a], ar[a], bi);
                }
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    dmma884(cr[a], -ai[a], bi);
                    dmma884(ci[a], ai[a], br);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int64_t zo = r0 + panel_col(cur.I, cur.J, warp * 8 + 2 * t + h) * ldz;
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    Z[zo + a * 8] = make_double2(cr[a][h], ci[a][h]);
                    if (Z32) Z32[zo + a * 8] = make_float2((float)cr[a][h], (float)ci[a][h]);
                }
            }
        }
        stage = (stage + 1) % U_NST;
        advance(cur);
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Cholesky-QR step of a 64-column panel (used by K4): sums the partial Grams, factors the diagonally scaled
// G = R^H R in shared memory, writes R (upper triangular) and W = R^-1, which the update kernel applies (P <- P W).
// `fail` is raised when a scaled pivot drops below `piv_tol` (panel numerically rank deficient for a Gram-based
// factorisation): the caller then falls back to the Householder TSQR for that panel.
__global__ void __launch_bounds__(256, 1)
    panel_chol_kernel(const c128* __restrict__ Gpart, int gram_ctas, int nchunk, c128* __restrict__ Wout,
                      c128* __restrict__ Rout, int64_t ldr, int* __restrict__ flags, int* __restrict__ fail,
                      double piv_tol, const double* __restrict__ before2, double dep_tol2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* G = reinterpret_cast<c128*>(smem_raw);
    c128* Wm = G + JP * GLD;
    __shared__ double d[JP];
    __shared__ int bad;
    const int tid = threadIdx.x;
    if (tid == 0) bad = 0;
    // every gram CTA holds one partial of this single pair in slot 2c
    {
        constexpr int PER = JP * JP / 256;  // 16 elements per thread, the loads of one slot issued together
        double sx[PER], sy[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) sx[j] = sy[j] = 0.0;
        for (int c = 0; c < gram_ctas; ++c) {
            const c128* src = Gpart + (size_t)(2 * c) * (JP * JP);
            c128 v[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) v[j] = src[tid + 256 * j];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                sx[j] += v[j].x;
                sy[j] += v[j].y;
            }
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int e = tid + 256 * j;
            G[(e >> 6) * GLD + (e & 63)] = make_double2(sx[j], sy[j]);
        }
    }
    __syncthreads();
    if (tid < JP) {
        double g = G[tid * GLD + tid].x;
        if (!(g > 0.0)) {
            bad = 1;
            g = 1.0;
        }
        // a column that lost (numerically) all of its norm to the earlier columns is rounding noise: the caller must
        // complete it with a genuinely new direction (qr.cu: robust_panel), not normalise the noise
        if (before2 && !(g > dep_tol2 * before2[tid])) bad = 1;
        d[tid] = sqrt(g);
    }
    __syncthreads();
    for (int e = tid; e < JP * JP; e += 256) {
        int r = e & 63, c = e >> 6;
        double sc = 1.0 / (d[r] * d[c]);
        c128 v = G[c * GLD + r];
        G[c * GLD + r] = make_double2(v.x * sc, v.y * sc);
    }
    __syncthreads();
    // right-looking Cholesky G = R^H R, one barrier per column: the scaled row j of R goes to Wm (scratch) while the
    // trailing update reads the unscaled row j of G
    for (int j = 0; j < JP; ++j) {
        double piv = G[j * GLD + j].x;
        if (!(piv > piv_tol)) {
            if (tid == 0) bad = 1;
            piv = fmax(piv, piv_tol);
        }
        const double inv = 1.0 / piv, rs = rsqrt(piv);
        const int nt = JP - j - 1;
        for (int e = tid; e < nt * nt + JP; e += 256) {
            if (e < JP) {  // row j of R
                int c = e;
                c128 v = G[c * GLD + j];
                Wm[c * GLD + j] = (c < j) ? make_double2(0.0, 0.0)
                                          : (c == j ? make_double2(piv * rs, 0.0) : make_double2(v.x * rs, v.y * rs));
            } else {
                int ee = e - JP;
                int i = j + 1 + ee % nt, c = j + 1 + ee / nt;
                if (i > c) continue;
                c128 gji = G[i * GLD + j], gjc = G[c * GLD + j];
                c128 pr = cmul(cconj(gji), gjc);
                c128 v = G[c * GLD + i];
                G[c * GLD + i] = make_double2(v.x - pr.x * inv, v.y - pr.y * inv);
            }
        }
        __syncthreads();
    }
    // R = Rs D (undo the scaling), kept in G; written out
    for (int e = tid; e < JP * JP; e += 256) {
        int r = e & 63, c = e >> 6;
        c128 v = Wm[c * GLD + r];
        v = (r <= c) ? make_double2(v.x * d[c], v.y * d[c]) : make_double2(0.0, 0.0);
        G[c * GLD + r] = v;
        Rout[r + (int64_t)c * ldr] = v;
    }
    __syncthreads();
    // W = R^-1 by back substitution: 4 threads per column split the inner sums (lanes 4c..4c+3 of one warp)
    {
        const int c = tid >> 2, part = tid & 3;
        c128* w = Wm + c * GLD;  // column c of W (overwrites the scratch copy of R, which now lives in G)
        for (int i = JP - 1; i >= 0; --i) {
            c128 acc = make_double2(0.0, 0.0);
            if (i < c)
                for (int k = i + 1 + part; k <= c; k += 4) acc = cadd(acc, cmul(G[k * GLD + i], w[k]));
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 1);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 1);
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, 2);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, 2);
            if (part == 0) {
                c128 v;
                if (i > c)
                    v = make_double2(0.0, 0.0);
                else {
                    double rii = 1.0 / G[i * GLD + i].x;  // real positive diagonal
                    v = (i == c) ? make_double2(rii, 0.0) : make_double2(-acc.x * rii, -acc.y * rii);
                }
                w[i] = v;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    for (int e = tid; e < JP * JP; e += 256) Wout[e] = Wm[(e >> 6) * GLD + (e & 63)];
    __syncthreads();
    if (tid == 0) {
        flags[0] = 1;
        if (bad) *fail = 1;
    }
}

// ---------------------------------------------------------------------------------------------
// set-up / finish kernels
// Z[:, :] = [A or A^H zero-padded ; I]
__global__ void svd_init_kernel(c128* __restrict__ Z, int64_t ldz, int mp, int np, int64_t m, int64_t n,
                                const c128* __restrict__ A, int64_t lda, int transposed) {
    int64_t total = ldz * np;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx % ldz, c = idx / ldz;
        c128 v = make_double2(0.0, 0.0);
        if (r < mp) {
            if (r < m && c < n) {
                if (!transposed)
                    v = A[r + c * lda];
                else {
                    c128 a = A[c + r * lda];
                    v = make_double2(a.x, -a.y);
                }
            }
        } else if (r - mp == c) {
            v.x = 1.0;
        }
        Z[idx] = v;
    }
}

// squared row and column norms of A (m x n, lda): one thread per row / per column
__global__ void rowcol_norm2_kernel(const c128* __restrict__ A, int64_t lda, int64_t m, int64_t n,
                                    double* __restrict__ rown, double* __restrict__ coln) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < m) {
        double acc = 0.0;
        for (int64_t j = 0; j < n; ++j) {
            c128 v = A[idx + j * lda];
            acc += v.x * v.x + v.y * v.y;
        }
        rown[idx] = acc;
    } else if (idx < m + n) {
        int64_t j = idx - m;
        double acc = 0.0;
        for (int64_t i = 0; i < m; ++i) {
            c128 v = A[i + j * lda];
            acc += v.x * v.x + v.y * v.y;
        }
        coln[j] = acc;
    }
}

// one warp per column: sigma_j = |X[:, j]|
__global__ void col_norm_kernel(const c128* __restrict__ Z, int64_t ldz, int mp, int np, double* __restrict__ sigma) {
    int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (col >= np) return;
    int lane = threadIdx.x & 31;
    const c128* x = Z + (int64_t)col * ldz;
    // scaled accumulation is unnecessary here: entries are O(sigma_1); plain sum of squares in fp64
    double acc = 0.0;
    for (int r = lane; r < mp; r += 32) {
        c128 v = x[r];
        acc += v.x * v.x + v.y * v.y;
    }
    acc = warp_sum(acc);
    if (lane == 0) sigma[col] = sqrt(acc);
}

// scale[0] = scale[1] = max_j sigma[j] (bit pattern of a non-negative double orders like an integer)
__global__ void set_double_kernel(double* p, double v) { *p = v; }

__global__ void max_reduce_kernel(const double* __restrict__ sigma, int n, double* __restrict__ scale) {
    __shared__ double sh[32];
    double mx = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, sigma[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) mx = fmax(mx, sh[i]);
        scale[0] = mx;
        scale[1] = mx;
        scale[2] = 0.0;  // absolute floor of the orthogonality test: off (pure relative criterion)
    }
}

// dst(i, j) or dst(j, i) = f(src(i, perm[j])) for i < rows, j < kept
// (output row = rowmap[i] when rowmap is given)
__global__ void svd_emit_kernel(const c128* __restrict__ src, int64_t ldz, int64_t rows, int64_t kept,
                                const int* __restrict__ perm, const double* __restrict__ sigma, int normalize,
                                int conj, int transpose_out, const int* __restrict__ rowmap, c128* __restrict__ dst,
                                int64_t ldd, const double* __restrict__ sc, int64_t sc_mod, int64_t sc_div) {
    int64_t total = rows * kept;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = idx % rows, j = idx / rows;
        int pj = perm[j];
        c128 v = src[i + (int64_t)pj * ldz];
        if (rowmap) i = rowmap[i];
        double f = 1.0;
        if (normalize) {
            double s = sigma[pj];
            f = (s > 0.0) ? 1.0 / s : 0.0;
        }
        if (sc) f *= sc[(sc_mod > 0 ? i % sc_mod : i) / sc_div];
        v.x *= f;
        v.y *= conj ? -f : f;
        if (transpose_out)
            dst[j + i * ldd] = v;
        else
            dst[i + j * ldd] = v;
    }
}

__global__ void svd_emit_sigma_kernel(const double* __restrict__ sigma, const int* __restrict__ perm, int64_t kept,
                                      double scale, double* __restrict__ S) {
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < kept) {
        double v = sigma[perm[j]];
        S[j] = (scale >= 0.0) ? v * scale : (v > 0.0 ? 1.0 / v : 0.0);  // scale < 0: reciprocal (0 for sigma = 0)
    }
}

// out(:, j) = Q(:, j) * r_jj / |r_jj|: undoes the column phases a Householder QR introduces
__global__ void fix_phase_kernel(const c128* __restrict__ Q, int64_t rows, int64_t cols, const c128* __restrict__ R,
                                 int64_t ldr, c128* __restrict__ out) {
    int64_t total = rows * cols;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = idx / rows;
        c128 d = R[j + j * ldr];
        double a = sqrt(d.x * d.x + d.y * d.y);
        c128 ph = (a > 0.0) ? make_double2(d.x / a, d.y / a) : make_double2(1.0, 0.0);
        out[idx] = cmul(Q[idx], ph);
    }
}

}  // namespace

struct SvdState {
    int64_t m, n, k;    // original problem, k = min(m, n)
    bool tall;          // m >= n: B = A, else B = A^H; B (rb x k) P = Q R, Jacobi runs on X0 = R^H (k x k)
    int64_t rb;         // rows of B
    int mp, np;         // padded k (rows of X, columns)
    int64_t ldz;
    c128* Z = nullptr;  // X (mp x np); the rotations are NOT accumulated (see qb_svd_emit)
    c128* B0 = nullptr;  // rb x k: B with its columns sorted by norm (B0 = Q R; Q itself is never needed)
    float2* Z32 = nullptr;  // FP32 shadow of X for the low-precision Gram of the early sweeps (null: not used)
    std::vector<double> sigma_sorted;
    double* sigma_dev = nullptr;
    int* perm_dev = nullptr;     // sigma order (descending) -> column of Z
    int* colperm_dev = nullptr;  // column j of the QR input was column colperm[j] of B
    std::vector<int> perm;
};

// QB200_UPDATE_3M=0 selects the 4-DMMA complex product in the update kernel (default: 3M, see the kernel)
static bool update_3m() {
    static const bool on = [] {
        const char* e = getenv("QB200_UPDATE_3M");
        return !(e && e[0] == '0');
    }();
    return on;
}

// QB200_EVD_WREG=0: the cross-rotation steps keep W in shared memory like the full steps (A/B switch; default: registers)
static bool evd_wreg() {
    static const bool on = [] {
        const char* e = getenv("QB200_EVD_WREG");
        return !(e && e[0] == '0');
    }();
    return on;
}

// Device-side sweep control of the Jacobi iteration (round 2): the whole iteration is ONE launch of a CUDA graph whose
// only node is a WHILE conditional; its body is the captured sweep followed by this one-thread kernel, which does what
// the host does between two sweeps -- read the largest coupling of the sweep, decide convergence, switch the absolute
// floor on for numerically rank-deficient input, carry the column-norm scale over -- and tells the graph whether to run
// the body again: no host round trip per sweep (11 per bulk TEBD bond otherwise).
struct SweepCtl {
    int sweeps, converged, floor_on, pad;
    double worst[48];
};
__global__ void jacobi_sweep_ctl_kernel(cudaGraphConditionalHandle handle, unsigned long long* __restrict__ stat,
                                        double* __restrict__ scale, SweepCtl* __restrict__ ctl, double conv_tol,
                                        int max_sweeps, int floor_after, double abs_floor) {
    const double worst = __longlong_as_double((long long)*stat);
    const int s = ctl->sweeps;  // index of the sweep that just ran
    if (s < 48) ctl->worst[s] = worst;
    ctl->sweeps = s + 1;
    *stat = 0ull;
    scale[0] = scale[1];
    const bool conv = !(worst > conv_tol);
    if (conv) {
        ctl->converged = 1;
    } else if (!ctl->floor_on && s >= floor_after) {
        ctl->floor_on = 1;
        scale[2] = abs_floor;
    }
    cudaGraphSetConditional(handle, (!conv && s + 1 < max_sweeps) ? 1u : 0u);
}

// QB200_SVD_WHILE=1 selects it.  Default OFF: measured on B200 (gpurun_out/r3i, same box, back to back) the sweep is not
// faster -- 3.19 s against 3.05 s per TEBD sweep for the best steps, identical sweep counts and results: with 12 bonds in
// flight the per-sweep read-back of one bond is hidden behind the kernels of the others, while the conditional graph
// costs more to instantiate per SVD.  Kept as the host-free form of the iteration (single-stream latency paths).
static bool while_enabled() {
    static const bool on = [] {
        const char* e = getenv("QB200_SVD_WHILE");
        return e && e[0] == '1';
    }();
    return on;
}

// QB200_SVD_MIXED=1: mixed-precision Jacobi (FP32 stage A + FP64 stage B, see qb_svd_factor)
static bool mixed_enabled() {
    static const bool on = [] {
        const char* e = getenv("QB200_SVD_MIXED");
        return e && e[0] == '1';
    }();
    return on;
}

// QB200_GRAM_LOWP=1 switches the mixed-precision Gram on (TF32 Gram while couplings > LOWP_TOL).  Default OFF:
// measured on B200 at k = 2048 (profiles/r1b_lowp_gram.txt) the Gram phase drops 19.5 -> 15.0 ms per bond (the
// per-launch cost is ramp / flush / tail, not DMMA) while the FP32 shadow stores cost the update kernel 3.3 ms:
// 1.8 % on the sweep, not worth a second code path in the convergence loop.  Kept for the experiment.
static bool gram_lowp() {
    static const bool on = [] {
        const char* e = getenv("QB200_GRAM_LOWP");
        return e && e[0] == '1';
    }();
    return on;
}

static unsigned grid_cap(qb200_ctx* ctx, int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads, cap = (int64_t)ctx->sm_count * 16;
    return (unsigned)std::max<int64_t>(1, std::min(b, cap));
}

void qb_svd_release(qb200_ctx* ctx, SvdState* st) {
    if (!st) return;
    if (st->Z) cudaFreeAsync(st->Z, ctx->stream);
    if (st->B0) cudaFreeAsync(st->B0, ctx->stream);
    if (st->Z32) cudaFreeAsync(st->Z32, ctx->stream);
    if (st->sigma_dev) cudaFreeAsync(st->sigma_dev, ctx->stream);
    if (st->perm_dev) cudaFreeAsync(st->perm_dev, ctx->stream);
    if (st->colperm_dev) cudaFreeAsync(st->colperm_dev, ctx->stream);
    delete st;
}

namespace {
// B0(:, j) = B(:, colperm[j]) with B = A (tall) or A^H (wide); B0 is rb x k, ld = rb
__global__ void svd_gather_cols_kernel(const c128* __restrict__ A, int64_t lda, int tall, int64_t rb, int64_t k,
                                       const int* __restrict__ colperm, c128* __restrict__ B0) {
    int64_t total = rb * k;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = idx % rb, j = idx / rb;
        int64_t c = colperm[j];
        c128 v;
        if (tall)
            v = A[i + c * lda];
        else {
            v = A[c + i * lda];  // B(i, c) = conj(A(c, i))
            v.y = -v.y;
        }
        B0[idx] = v;
    }
}
}  // namespace

int32_t qb_svd_init(qb200_ctx* ctx) {
    QB_CUDA(ctx, cudaFuncSetAttribute(jacobi_gram_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(jacobi_gram_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(jacobi_evd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EVD_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(jacobi_evd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EVD_SMEM_WREG));
    QB_CUDA(ctx, cudaFuncSetAttribute(jacobi_gram32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GRAM32_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(jacobi_update_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(jacobi_update_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(panel_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)((size_t)2 * JP * GLD * sizeof(c128))));
    return QB200_OK;
}

int32_t qb_svd_factor(qb200_ctx* ctx, int64_t m, int64_t n, const c128* A, int64_t lda, SvdState** out,
                      std::vector<double>& sigma) {
    if (m <= 0 || n <= 0) QB_FAIL(ctx, QB200_E_INVALID, "svd: empty matrix");
    if (std::max(m, n) > (1 << 30) || std::min(m, n) > 16384)
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "svd: matrix too large (min(m, n) <= 16384)");
    SvdState* st = new SvdState();
    st->m = m;
    st->n = n;
    st->tall = m >= n;
    st->k = std::min(m, n);
    st->rb = std::max(m, n);
    const int64_t k = st->k, rb = st->rb;
    st->mp = st->np = (int)((k + 63) / 64 * 64);
    st->ldz = (int64_t)st->mp;
    const int nb = st->np / JB, npairs = nb / 2, nsteps = (nb == 2) ? 1 : nb - 1;
    auto fail = [&](int32_t code) {
        qb_svd_release(ctx, st);
        return code;
    };
    auto cuda_fail = [&](cudaError_t e) {
        ctx->err = std::string("svd: ") + cudaGetErrorString(e);
        return fail(QB200_E_CUDA);
    };
    if (cudaMallocAsync(&st->Z, sizeof(c128) * st->ldz * st->np, ctx->stream) != cudaSuccess ||
        cudaMallocAsync(&st->B0, sizeof(c128) * rb * k, ctx->stream) != cudaSuccess ||
        cudaMallocAsync(&st->sigma_dev, sizeof(double) * st->np, ctx->stream) != cudaSuccess ||
        cudaMallocAsync(&st->perm_dev, sizeof(int) * st->np, ctx->stream) != cudaSuccess ||
        cudaMallocAsync(&st->colperm_dev, sizeof(int) * k, ctx->stream) != cudaSuccess) {
        ctx->err = "svd: out of device memory";
        return fail(QB200_E_CUDA);
    }
    Workspace ws(ctx);
    // ---- preconditioner (Drmac & Veselic): B P = Q R with the columns of B = A or A^H sorted by decreasing norm,
    //      then one-sided Jacobi on X0 = R^H.  R^H is column graded with a well-conditioned scaled part whatever
    //      the grading of A (row, column or both, as for theta = Λl Γ Λ Γ Λr), which both halves the number of
    //      sweeps on graded matrices and gives the small singular vectors relative accuracy. ----
    {
        double* nr = ws.get<double>((size_t)(m + n));
        c128* B0 = st->B0;
        Workspace wsq(ctx);  // Q and R only live until X = R^H is set up
        c128* Qtmp = wsq.get<c128>((size_t)(rb * k));
        c128* R = wsq.get<c128>((size_t)(k * k));
        if (!nr || !Qtmp || !R) {
            ctx->err = "svd: workspace allocation failed";
            return fail(QB200_E_CUDA);
        }
        rowcol_norm2_kernel<<<(unsigned)((m + n + 127) / 128), 128, 0, ctx->stream>>>(A, lda, m, n, nr, nr + m);
        ctx->launches++;
        std::vector<double> h((size_t)(m + n));
        cudaError_t e = cudaMemcpyAsync(h.data(), nr, sizeof(double) * (m + n), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = qb_stream_sync(ctx);
        if (e != cudaSuccess) return cuda_fail(e);
        const double* cn = st->tall ? h.data() + m : h.data();  // column norms of B
        std::vector<int> cp((size_t)k);
        std::iota(cp.begin(), cp.end(), 0);
        std::stable_sort(cp.begin(), cp.end(), [&](int a, int b) { return cn[a] > cn[b]; });
        e = cudaMemcpyAsync(st->colperm_dev, cp.data(), sizeof(int) * k, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = qb_stream_sync(ctx);
        if (e != cudaSuccess) return cuda_fail(e);
        svd_gather_cols_kernel<<<grid_cap(ctx, rb * k, 256), 256, 0, ctx->stream>>>(A, lda, st->tall ? 1 : 0, rb, k,
                                                                                   st->colperm_dev, B0);
        ctx->launches++;
        {
            PhaseTimer pt(ctx, QB_PH_QR, 8.0 * (2.0 * rb * k * k - 2.0 / 3.0 * k * k * k));
            // only R is used afterwards (U S = A V, see qb_svd_emit).  Both Gram-Schmidt passes are still needed:
            // with one pass the singular values of R deviate from those of A by the loss of orthogonality of Q,
            // ~ kappa(A) eps (measured 1e-10 sigma_1 on a 1e10-graded matrix), R = R2 R1 restores eps sigma_1.
            int32_t r = qb_qr_matrix(ctx, rb, k, B0, rb, Qtmp, rb, R, k, QB_QR_R_ONLY);
            if (r != QB200_OK) return fail(r);
        }
        // Z = R^H zero-padded
        svd_init_kernel<<<grid_cap(ctx, st->ldz * st->np, 256), 256, 0, ctx->stream>>>(st->Z, st->ldz, st->mp, st->np, k,
                                                                                      k, R, k, 1);
        ctx->launches++;
    }

    // persistent grids: gram and update 2 CTAs per SM, never more CTAs than work items
    const int g_nchunk = st->mp / G_BKR, u_nchunk = (int)(st->ldz / U_ROWS);
    const int gram_ctas = std::max(1, std::min(2 * ctx->sm_count, npairs * g_nchunk));
    const int upd_ctas = std::max(1, std::min(2 * ctx->sm_count, npairs * u_nchunk));

    c128* Gpart = ws.get<c128>((size_t)2 * gram_ctas * JP * JP);
    c128* Wg = ws.get<c128>((size_t)npairs * JP * JP);
    int* flags = ws.get<int>(npairs);
    unsigned long long* stat = ws.get<unsigned long long>(1);
    double* scale = ws.get<double>(3);  // [0] largest column norm seen in the previous sweep, [1] running max, [2] absolute floor factor
    c128* Dstore = (nb > 2) ? ws.get<c128>((size_t)nb * JB * JB) : nullptr;  // carried diagonal Gram blocks
    if (!Gpart || !Wg || !flags || !stat || !scale || (nb > 2 && !Dstore)) {
        ctx->err = "svd: workspace allocation failed";
        return fail(QB200_E_CUDA);
    }

    const double eps = 1.1102230246251565e-16;
    const double rot_tol = std::sqrt((double)st->mp) * eps;
    const double conv_tol = 1e-7;  // quadratic convergence: what is left after such a sweep is <~ 40 worst^2
    // Orthogonality test of a column pair: |g_ij| <= rot_tol |x_i||x_j| (pure relative: R^H is column graded, so one-sided
    // Jacobi keeps relative accuracy and there is no noise floor) -- unless the matrix is numerically RANK DEFICIENT.  Then
    // some columns of X are rounding noise (norm ~ eps sigma_1 sqrt(updates)), every GEMM-applied rotation regenerates
    // their O(1) relative couplings and the relative test never passes (measured: 40 sweeps without convergence on the
    // 5120 x 2560, rank-1536 site matrix of an MPO-applied state).  For such inputs -- the QR preconditioner reports
    // dependent columns, or the iteration is still running after 14 sweeps -- pairs also count as orthogonal when
    // |g_ij| <= ABS_FLOOR sigma_1 max(|x_i|, |x_j|): columns below the floor are left alone.  Singular values keep their
    // absolute accuracy eps sigma_1 (the bar is 1e-12 sigma_1); only values below 1e-13 sigma_1 lose relative digits.
    const double abs_c = -1.0;  // read from scale[2]
    const double ABS_FLOOR = 1024.0 * 1.1102230246251565e-16;
    bool floor_on = ctx->qr_last_dependent > 0;
    const int inner_sweeps = (nb == 2) ? 12 : 1;
    const int nact = (nb == 2) ? (int)std::min<int64_t>(64, (k + 1) / 2 * 2) : 64;
    col_norm_kernel<<<(st->np + 7) / 8, 256, 0, ctx->stream>>>(st->Z, st->ldz, st->mp, st->np, st->sigma_dev);
    max_reduce_kernel<<<1, 256, 0, ctx->stream>>>(st->sigma_dev, st->np, scale);
    ctx->launches += 2;
    if (floor_on) {
        set_double_kernel<<<1, 1, 0, ctx->stream>>>(scale + 2, ABS_FLOOR);
        ctx->launches++;
    }
    // mixed-precision Gram (see jacobi_gram32_kernel): sweep 0 runs in FP64 and measures the couplings; while the
    // largest coupling of the previous sweep is above LOWP_TOL the cross Grams come from the FP32 shadow of X, which
    // the update kernel keeps in step.  Once below, the iteration is FP64 only for good.
    const double LOWP_TOL = 1e-2;
    bool shadow = gram_lowp() && nb >= 8, lowp = false;
    if (shadow) {
        if (cudaMallocAsync(&st->Z32, sizeof(float2) * st->ldz * st->np, ctx->stream) != cudaSuccess) {
            ctx->err = "svd: out of device memory";
            return fail(QB200_E_CUDA);
        }
        int32_t r = qb_narrow_c128(ctx, st->Z, st->Z32, st->ldz * st->np);
        if (r != QB200_OK) return fail(r);
    }
    // ---- mixed precision: the linear phase of the iteration on an FP32 shadow ("stage A") ----
    // The first ~8 of ~11 sweeps only steer: their rotations need no FP64 accuracy, but X <- X W_p applied in FP64 is
    // 80 % of the SVD's tensor-pipe time.  Stage A runs the same block iteration on S = [X32; V32] (FP32 image of X
    // and the accumulated rotations) -- cross Grams on the TF32 path, rotations by the same evd kernel, updates on the
    // FP32 / TF32 tensor path -- until the couplings reach the FP32 noise level.  Then V32 is widened, orthonormalised
    // in FP64 (two Newton-Schulz steps: its unitarity defect ~1e-6 -> 1e-12 -> 1e-24), X <- X0 V is ONE FP64 GEMM, and
    // the FP64 iteration ("stage B", the loop below, unchanged) starts from nearly orthogonal columns: 2-3 sweeps.
    // The result is an exact unitary image of X0 whatever stage A did; only the sweep count depends on it.
    int sweeps_a = 0;
    if (mixed_enabled() && nb >= 16 && !floor_on && !shadow) {
        Workspace wsa(ctx);
        const int64_t lds = (int64_t)2 * st->mp;
        const int np = st->np, mp = st->mp;
        float2* S = wsa.get<float2>((size_t)lds * np);
        int nsplit = 1;  // largest divisor of mp / 32 that is <= 8: every split sums a whole number of 32-row tiles
        for (int d = 2; d <= 8; ++d)
            if ((mp / 32) % d == 0) nsplit = d;
        c128* GpartA = wsa.get<c128>((size_t)2 * std::max(gram_ctas, npairs * nsplit) * JP * JP);
        if (!S || !GpartA) {
            ctx->err = "svd: workspace allocation failed";
            return fail(QB200_E_CUDA);
        }
        lp_init_kernel<<<grid_cap(ctx, lds * np, 256), 256, 0, ctx->stream>>>(st->Z, st->ldz, mp, np, S, lds);
        ctx->launches++;
        static const int lp_mode = [] {  // QB200_LP_UPDATE: (default) tcgen05 kernel, "simple" FP32 FMA kernel, "check" both
            const char* e = getenv("QB200_LP_UPDATE");
            return !e ? 0 : (e[0] == 's' ? 1 : (e[0] == 'c' ? 2 : 0));
        }();
        const double rot_tol_a = 2e-6, tol_a = 3e-4;
        const int max_a = 14;
        double prev = 1e300;
        int32_t emit_rc = QB200_OK;
        auto emit_sweep_a = [&]() {
            cudaMemsetAsync(stat, 0, sizeof(unsigned long long), ctx->stream);
            for (int step = 0; step < nsteps; ++step) {
                const int mode = (step == 0) ? 1 : 2;
                {
                    PhaseTimer pt(ctx, QB_PH_LP_GRAM, 8.0 * npairs * (double)mp * JP * JP);
                    if (mode == 1)
                        lp_gram_full_kernel<<<dim3(nsplit, npairs), 256, 0, ctx->stream>>>(S, lds, mp / nsplit, nb, step, GpartA);
                    else
                        jacobi_gram32_kernel<<<gram_ctas, 256, GRAM32_SMEM, ctx->stream>>>(S, lds, mp, nb, step, npairs, GpartA);
                }
                {
                    PhaseTimer pt(ctx, QB_PH_JEVD, 0.0);
                    (mode == 2 && evd_wreg() ? jacobi_evd_kernel<1> : jacobi_evd_kernel<0>)<<<npairs, EVD_THREADS, mode == 2 && evd_wreg() ? EVD_SMEM_WREG : EVD_SMEM, ctx->stream>>>(
                        GpartA, mode == 1 ? npairs * nsplit : gram_ctas, mode == 1 ? nsplit : g_nchunk, npairs, Wg, flags, stat,
                        rot_tol_a, inner_sweeps, scale, (unsigned long long*)(scale + 1), abs_c, nact, mode, Dstore, nb, step);
                }
                if (lp_mode == 2 && sweeps_a == 0 && step < 2) {  // QB200_LP_UPDATE=check: tcgen05 kernel against the FP32 FMA one
                    float2* T = nullptr;
                    unsigned* dmax = nullptr;
                    cudaMallocAsync(&T, sizeof(float2) * lds * np, ctx->stream);
                    cudaMallocAsync(&dmax, 2 * sizeof(unsigned), ctx->stream);
                    cudaMemsetAsync(dmax, 0, 2 * sizeof(unsigned), ctx->stream);
                    cudaMemcpyAsync(T, S, sizeof(float2) * lds * np, cudaMemcpyDeviceToDevice, ctx->stream);
                    lp_update_simple_kernel<<<dim3((unsigned)(lds / 32), npairs), 256, 0, ctx->stream>>>(T, lds, nb, step, Wg, flags);
                    int32_t r = launch_lp_update_tc5(ctx, S, lds, lds, nb, step, Wg, flags, npairs);
                    if (r != QB200_OK) emit_rc = r;
                    lp_maxdiff_kernel<<<grid_cap(ctx, lds * np, 256), 256, 0, ctx->stream>>>(S, T, lds * np, dmax);
                    unsigned h[2];
                    cudaMemcpyAsync(h, dmax, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream);
                    cudaStreamSynchronize(ctx->stream);
                    float fd, fm;
                    memcpy(&fd, &h[0], 4);
                    memcpy(&fm, &h[1], 4);
                    fprintf(stderr, "[qb200 svd] lp update check step %d: max |tc5 - fma| = %.3e, max |fma| = %.3e\n", step, fd, fm);
                    cudaFreeAsync(T, ctx->stream);
                    cudaFreeAsync(dmax, ctx->stream);
                } else {
                    PhaseTimer pt(ctx, QB_PH_LP_UPDATE, 8.0 * npairs * (double)lds * JP * JP);
                    if (lp_mode == 1) {
                        lp_update_simple_kernel<<<dim3((unsigned)(lds / 32), npairs), 256, 0, ctx->stream>>>(S, lds, nb, step, Wg, flags);
                    } else {
                        int32_t r = launch_lp_update_tc5(ctx, S, lds, lds, nb, step, Wg, flags, npairs);
                        if (r != QB200_OK) emit_rc = r;
                    }
                }
            }
            cudaMemcpyAsync(scale, scale + 1, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
        };
        // one stage-A sweep = 3 (nb - 1) launches with sweep-independent arguments: captured once, replayed
        cudaGraphExec_t graph_a = nullptr;
        static const bool graph_a_enabled = [] {
            const char* e = getenv("QB200_SVD_GRAPH");
            return !(e && e[0] == '0');
        }();
        if (graph_a_enabled && lp_mode == 0 && !ctx->prof_on && !qb_sync_debug()) {
            cudaGraph_t g = nullptr;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                emit_sweep_a();
                cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
                if (ce == cudaSuccess && g && emit_rc == QB200_OK && cudaGraphInstantiate(&graph_a, g, 0) != cudaSuccess) graph_a = nullptr;
                if (g) cudaGraphDestroy(g);
                if (!graph_a) cudaGetLastError();  // fall back to plain launches
                emit_rc = QB200_OK;
            }
        }
        for (; sweeps_a < max_a; ) {
            if (graph_a)
                cudaGraphLaunch(graph_a, ctx->stream);
            else
                emit_sweep_a();
            if (emit_rc != QB200_OK) {
                if (graph_a) cudaGraphExecDestroy(graph_a);
                return fail(emit_rc);
            }
            ctx->launches += 3 * (int64_t)nsteps;
            ++sweeps_a;
            cudaError_t e = cudaMemcpyAsync(ctx->scratch_host, stat, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = qb_stream_sync(ctx);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess) {
                if (graph_a) cudaGraphExecDestroy(graph_a);
                return cuda_fail(e);
            }
            const double worst = ctx->scratch_host[0];
            if (getenv("QB200_DEBUG"))
                fprintf(stderr, "[qb200 svd] stage A (fp32 shadow) sweep %d worst %.3e\n", sweeps_a - 1, worst);
            if (!(worst > tol_a)) break;
            if (prev < 1e-2 && worst > 0.3 * prev) break;  // at the noise floor of the shadow
            prev = worst;
        }
        if (graph_a) cudaGraphExecDestroy(graph_a);
        {
            PhaseTimer pt(ctx, QB_PH_LP_GLUE, 8.0 * (5.0 * np * (double)np * np));
            c128* Va = wsa.get<c128>((size_t)np * np);
            c128* Vb = wsa.get<c128>((size_t)np * np);
            c128* T = wsa.get<c128>((size_t)np * np);
            c128* X1 = nullptr;
            if (!Va || !Vb || !T || cudaMallocAsync(&X1, sizeof(c128) * st->ldz * st->np, ctx->stream) != cudaSuccess) {
                ctx->err = "svd: out of device memory";
                return fail(QB200_E_CUDA);
            }
            lp_extract_v_kernel<<<grid_cap(ctx, (int64_t)np * np, 256), 256, 0, ctx->stream>>>(S, lds, mp, np, Va);
            ctx->launches++;
            const c128 one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0), mhalf = make_double2(-0.5, 0.0);
            int32_t r = QB200_OK;
            for (int it = 0; it < 2 && r == QB200_OK; ++it) {  // V <- V (3/2 I - 1/2 V^H V)
                r = qb_gemm(ctx, 2, 0, np, np, np, mhalf, Va, np, Va, np, zero, T, np);
                if (r == QB200_OK) {
                    add_diag_kernel<<<(np + 255) / 256, 256, 0, ctx->stream>>>(T, np, 1.5);
                    ctx->launches++;
                    r = qb_gemm(ctx, 0, 0, np, np, np, one, Va, np, T, np, zero, Vb, np);
                }
                std::swap(Va, Vb);
            }
            if (r == QB200_OK) r = qb_gemm(ctx, 0, 0, mp, np, np, one, st->Z, st->ldz, Va, np, zero, X1, st->ldz);
            if (r != QB200_OK) {
                cudaFreeAsync(X1, ctx->stream);
                return fail(r);
            }
            cudaFreeAsync(st->Z, ctx->stream);
            st->Z = X1;
        }
    }

    const int max_sweeps = 40;
    int sweep = 0;
    bool converged = false;
    // One Jacobi sweep = 3 kernels x (nb - 1) steps whose arguments are the same in every sweep: the sweep is captured
    // ONCE per SVD as a CUDA graph and replayed (189 launches -> 1 graph launch per sweep at k = 2048; the host thread of
    // a worker stream issues ~10x fewer driver calls per SVD).  Plain launches with the phase profiler, the
    // mixed-precision Gram experiment, QB200_SYNC_DEBUG, QB200_SVD_GRAPH=0 and for single-pair problems.
    auto emit_sweep = [&](bool lowp_now, bool kernels_only = false) {  // kernels_only: the WHILE body (its control kernel
                                                                       // resets stat and carries the scale over)
        if (!kernels_only) cudaMemsetAsync(stat, 0, sizeof(unsigned long long), ctx->stream);
        for (int step = 0; step < nsteps; ++step) {
            const int mode = (nb == 2) ? 0 : (step == 0 ? 1 : 2);
            {
                PhaseTimer pt(ctx, QB_PH_JGRAM, 8.0 * npairs * (double)st->mp * JP * JP);  // full-product count
                if (mode == 2 && lowp_now)
                    jacobi_gram32_kernel<<<gram_ctas, 256, GRAM32_SMEM, ctx->stream>>>(st->Z32, st->ldz, st->mp, nb, step,
                                                                                      npairs, Gpart);
                else if (mode == 2)
                    jacobi_gram_kernel<1><<<gram_ctas, 256, GRAM_SMEM, ctx->stream>>>(st->Z, st->ldz, st->mp, nb, step,
                                                                                     npairs, Gpart);
                else
                    jacobi_gram_kernel<0><<<gram_ctas, 256, GRAM_SMEM, ctx->stream>>>(st->Z, st->ldz, st->mp, nb, step,
                                                                                     npairs, Gpart);
            }
            {
                PhaseTimer pt(ctx, QB_PH_JEVD, 0.0);
                (mode == 2 && evd_wreg() ? jacobi_evd_kernel<1> : jacobi_evd_kernel<0>)<<<npairs, EVD_THREADS, mode == 2 && evd_wreg() ? EVD_SMEM_WREG : EVD_SMEM, ctx->stream>>>(
                    Gpart, gram_ctas, g_nchunk, npairs, Wg, flags, stat, rot_tol, inner_sweeps, scale,
                    (unsigned long long*)(scale + 1), abs_c, nact, mode, Dstore, nb, step);
            }
            {
                PhaseTimer pt(ctx, QB_PH_JUPDATE, 8.0 * npairs * (double)st->ldz * JP * JP);  // ldz = rows of X
                (update_3m() ? jacobi_update_kernel<1> : jacobi_update_kernel<0>)<<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(
                    st->Z, st->ldz, nb, step, Wg, flags, npairs, u_nchunk, shadow ? st->Z32 : nullptr);
            }
        }
        if (!kernels_only) cudaMemcpyAsync(scale, scale + 1, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
    };
    static const bool graph_enabled = [] {
        const char* e = getenv("QB200_SVD_GRAPH");
        return !(e && e[0] == '0');
    }();
    const bool graph_ok = graph_enabled && nb > 2 && !ctx->prof_on && !shadow && !qb_sync_debug();
    // ---- the whole iteration as one WHILE graph (device-side convergence) ----
    bool while_done = false;
    if (graph_ok && while_enabled()) {
        SweepCtl* ctl = ws.get<SweepCtl>(1);
        cudaGraph_t g = nullptr;
        cudaGraphExec_t gx = nullptr;
        bool built = false;
        if (ctl && cudaGraphCreate(&g, 0) == cudaSuccess) {
            cudaGraphConditionalHandle handle;
            cudaGraphNodeParams np_ = {};
            np_.type = cudaGraphNodeTypeConditional;
            cudaGraphNode_t node;
            if (cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault) == cudaSuccess) {
                np_.conditional.handle = handle;
                np_.conditional.type = cudaGraphCondTypeWhile;
                np_.conditional.size = 1;
                if (cudaGraphAddNode(&node, g, nullptr, 0, &np_) == cudaSuccess &&
                    cudaStreamBeginCaptureToGraph(ctx->stream, np_.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                                  cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    emit_sweep(false, true);
                    jacobi_sweep_ctl_kernel<<<1, 1, 0, ctx->stream>>>(handle, stat, scale, ctl, conv_tol, max_sweeps, 13,
                                                                      ABS_FLOOR);
                    built = cudaStreamEndCapture(ctx->stream, nullptr) == cudaSuccess &&
                            cudaGraphInstantiate(&gx, g, 0) == cudaSuccess;
                }
            }
        }
        if (built) {
            SweepCtl init;
            memset(&init, 0, sizeof(init));
            init.floor_on = floor_on ? 1 : 0;
            SweepCtl* hctl = reinterpret_cast<SweepCtl*>(ctx->scratch_host + 16);  // pinned page; [0 .. 15] is used elsewhere
            *hctl = init;
            cudaError_t e = cudaMemcpyAsync(ctl, hctl, sizeof(SweepCtl), cudaMemcpyHostToDevice, ctx->stream);
            if (e == cudaSuccess) e = cudaMemsetAsync(stat, 0, sizeof(unsigned long long), ctx->stream);
            if (e == cudaSuccess) e = cudaGraphLaunch(gx, ctx->stream);
            if (e == cudaSuccess) e = cudaMemcpyAsync(hctl, ctl, sizeof(SweepCtl), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = qb_stream_sync(ctx);
            if (e == cudaSuccess) e = cudaGetLastError();
            cudaGraphExecDestroy(gx);
            cudaGraphDestroy(g);
            if (e != cudaSuccess) return cuda_fail(e);
            sweep = hctl->sweeps;
            converged = hctl->converged != 0;
            floor_on = hctl->floor_on != 0;
            ctx->launches += (3 * (int64_t)nsteps + 1) * sweep;
            if (getenv("QB200_DEBUG"))
                for (int i = 0; i < sweep && i < 48; ++i)
                    fprintf(stderr, "[qb200 svd] %lld x %lld (jacobi on %lld^2, nb %d) sweep %d worst %.3e\n", (long long)m,
                            (long long)n, (long long)k, nb, i, hctl->worst[i]);
            while_done = true;
        } else {
            if (gx) cudaGraphExecDestroy(gx);
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();  // fall back to one graph launch per sweep
        }
    }
    cudaGraphExec_t sweep_graph = nullptr;
    if (graph_ok && !while_done) {
        cudaGraph_t g = nullptr;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            emit_sweep(false);
            cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
            if (ce == cudaSuccess && g && cudaGraphInstantiate(&sweep_graph, g, 0) != cudaSuccess) sweep_graph = nullptr;
            if (g) cudaGraphDestroy(g);
            if (!sweep_graph) cudaGetLastError();  // fall back to plain launches
        }
    }
    for (; !while_done && sweep < max_sweeps && !converged; ++sweep) {
        if (sweep_graph)
            cudaGraphLaunch(sweep_graph, ctx->stream);
        else
            emit_sweep(lowp);
        ctx->launches += 3 * (int64_t)nsteps;
        cudaError_t e = cudaMemcpyAsync(ctx->scratch_host, stat, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = qb_stream_sync(ctx);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            if (sweep_graph) cudaGraphExecDestroy(sweep_graph);
            return cuda_fail(e);
        }
        double worst = ctx->scratch_host[0];
        if (getenv("QB200_DEBUG"))
            fprintf(stderr, "[qb200 svd] %lld x %lld (jacobi on %lld^2, nb %d) sweep %d%s worst %.3e\n", (long long)m,
                    (long long)n, (long long)k, nb, sweep, lowp ? " (tf32 gram)" : "", worst);
        if (!(worst > conv_tol)) converged = true;
        if (!converged && !floor_on && sweep >= 13) {  // numerically rank deficient after all: switch the floor on
            floor_on = true;
            set_double_kernel<<<1, 1, 0, ctx->stream>>>(scale + 2, ABS_FLOOR);
            ctx->launches++;
        }
        if (shadow) {
            lowp = worst > LOWP_TOL;
            if (!lowp) shadow = false;  // quadratic phase ahead: FP64 Gram from here on, the shadow is no longer kept
        }
    }
    if (sweep_graph) cudaGraphExecDestroy(sweep_graph);
    ctx->last_svd_sweeps = sweep + sweeps_a;
    ctx->svd_calls++;
    ctx->svd_sweeps += sweep + sweeps_a;
    if (!converged) {
        ctx->err = "svd: Jacobi did not converge within the sweep limit";
        return fail(QB200_E_NOCONVERGE);
    }
    col_norm_kernel<<<(st->np + 7) / 8, 256, 0, ctx->stream>>>(st->Z, st->ldz, st->mp, st->np, st->sigma_dev);
    ctx->launches++;
    std::vector<double> sig(st->np);
    cudaError_t e = cudaMemcpyAsync(sig.data(), st->sigma_dev, sizeof(double) * st->np, cudaMemcpyDeviceToHost,
                                    ctx->stream);
    if (e == cudaSuccess) e = qb_stream_sync(ctx);
    if (e != cudaSuccess) return cuda_fail(e);
    // descending, ties broken by the lower column index (first occurrence); padded columns come last
    st->perm.resize(st->np);
    std::iota(st->perm.begin(), st->perm.end(), 0);
    std::stable_sort(st->perm.begin(), st->perm.end(), [&](int a, int b) {
        bool pa = a >= k, pb = b >= k;
        if (pa != pb) return pb;
        return sig[a] > sig[b];
    });
    sigma.resize(k);
    for (int64_t i = 0; i < k; ++i) sigma[i] = sig[st->perm[i]];
    st->sigma_sorted = sigma;
    e = cudaMemcpyAsync(st->perm_dev, st->perm.data(), sizeof(int) * st->np, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = qb_stream_sync(ctx);
    if (e != cudaSuccess) return cuda_fail(e);
    *out = st;
    return QB200_OK;
}

// B P = Q R, R^H = Xn S Vs^H  =>  B = (Q Vs) S (P Xn)^H with Q Vs = B P Xn S^-1 =: Y.  tall: A = B, U = Y, V = P Xn.
// wide: A = B^H, U = P Xn, V = Y.  (Xn = normalised columns of X in sigma order; (P Xn)(colperm[j], :) = Xn(j, :).)
int32_t qb_svd_emit(qb200_ctx* ctx, SvdState* st, int64_t kept, c128* U, int64_t ldu, double* S, c128* V,
                    int64_t ldv, int vmode, const double* uinv, int64_t uinv_len, const double* vinv,
                    int64_t vinv_div, double sigma_scale) {
    if (kept <= 0) return QB200_OK;
    const int64_t k = st->k;
    const c128* X = st->Z;
    const c128 ONE = make_double2(1.0, 0.0), ZERO = make_double2(0.0, 0.0);
    Workspace ws(ctx);
    // Exactly zero singular values (zero columns of X: trailing dependent columns, zero input): their X-side vectors are
    // zero columns, not unit vectors.  Complete them to an orthonormal set with K4 (robust_panel replaces dependent
    // columns by new orthogonal directions) and emit from the completed factor instead of X.
    const c128* Xsrc = X;
    int64_t xld = st->ldz;
    const int* xperm = st->perm_dev;
    int xnorm = 1;
    if (st->sigma_sorted[kept - 1] == 0.0) {
        c128* Xn = ws.get<c128>((size_t)(k * kept));
        c128* Qx = ws.get<c128>((size_t)(k * kept));
        c128* Rx = ws.get<c128>((size_t)(kept * kept));
        int* ident = ws.get<int>((size_t)kept);
        if (!Xn || !Qx || !Rx || !ident) QB_FAIL(ctx, QB200_E_CUDA, "svd: workspace allocation failed");
        svd_emit_kernel<<<grid_cap(ctx, k * kept, 256), 256, 0, ctx->stream>>>(X, st->ldz, k, kept, st->perm_dev, st->sigma_dev,
                                                                               1, 0, 0, nullptr, Xn, k, nullptr, 0, 1);
        QB_LAUNCH_CHECK(ctx);
        QB_TRY(qb_qr_matrix(ctx, k, kept, Xn, k, Qx, k, Rx, kept, 2));
        fix_phase_kernel<<<grid_cap(ctx, k * kept, 256), 256, 0, ctx->stream>>>(Qx, k, kept, Rx, kept, Xn);
        QB_LAUNCH_CHECK(ctx);
        std::vector<int> id((size_t)kept);
        std::iota(id.begin(), id.end(), 0);
        QB_CUDA(ctx, cudaMemcpyAsync(ident, id.data(), sizeof(int) * kept, cudaMemcpyHostToDevice, ctx->stream));
        QB_CUDA(ctx, qb_stream_sync(ctx));
        Xsrc = Xn;
        xld = k;
        xperm = ident;
        xnorm = 0;
    }
    // Neither the rotations nor Q were kept.  With B0 = B P = Q R and R^H = Xn S Vs^H (Xn orthonormal):
    // B0 Xn = Q R Xn = Q Vs S, so the "Q side" factor is Y = B0 Xn S^-1 -- ONE GEMM with the (sorted) input matrix.
    // Column j of Y carries a relative error ~ eps sigma_1 / sigma_j: harmless (norm-wise backward stable), but it
    // costs orthogonality when the kept spectrum is graded; then the columns are re-orthonormalised in order of
    // decreasing sigma (K4, phases fixed), which keeps A = U S V^H to eps sigma_1 and Y orthonormal to eps.
    auto make_y = [&](c128* Y) -> int32_t {  // Y: rb x kept, ld = rb
        c128* Xsel = ws.get<c128>((size_t)(k * kept));
        double* isg = ws.get<double>((size_t)kept);
        if (!Xsel || !isg) QB_FAIL(ctx, QB200_E_CUDA, "svd: workspace allocation failed");
        svd_emit_kernel<<<grid_cap(ctx, k * kept, 256), 256, 0, ctx->stream>>>(
            Xsrc, xld, k, kept, xperm, st->sigma_dev, xnorm, 0, 0, nullptr, Xsel, k, nullptr, 0, 1);
        QB_LAUNCH_CHECK(ctx);
        svd_emit_sigma_kernel<<<(unsigned)((kept + 255) / 256), 256, 0, ctx->stream>>>(st->sigma_dev, st->perm_dev, kept,
                                                                                      -1.0, isg);
        QB_LAUNCH_CHECK(ctx);
        QB_TRY(qb_gemm(ctx, 0, 0, st->rb, kept, k, ONE, st->B0, st->rb, Xsel, k, ZERO, Y, st->rb));
        QB_TRY(qb_scale_rows_cols(ctx, Y, Y, st->rb, kept, nullptr, 1, isg, 1));
        const std::vector<double>& sg = st->sigma_sorted;
        bool graded = !(sg[kept - 1] > 1e-3 * sg[0]);
        if (graded) {
            c128* Qy = ws.get<c128>((size_t)(st->rb * kept));
            c128* Ry = ws.get<c128>((size_t)(kept * kept));
            if (!Qy || !Ry) QB_FAIL(ctx, QB200_E_CUDA, "svd: workspace allocation failed");
            QB_TRY(qb_qr_matrix(ctx, st->rb, kept, Y, st->rb, Qy, st->rb, Ry, kept, 2));
            fix_phase_kernel<<<grid_cap(ctx, st->rb * kept, 256), 256, 0, ctx->stream>>>(Qy, st->rb, kept, Ry, kept, Y);
            QB_LAUNCH_CHECK(ctx);
        }
        return QB200_OK;
    };
    PhaseTimer pt(ctx, QB_PH_EMIT, 8.0 * st->rb * k * kept);
    if (U) {
        if (st->tall) {  // U = Y, rows scaled by uinv[i % uinv_len]
            if (ldu == st->rb) {
                QB_TRY(make_y(U));
            } else {
                c128* Y = ws.get<c128>((size_t)(st->rb * kept));
                if (!Y) QB_FAIL(ctx, QB200_E_CUDA, "svd: workspace allocation failed");
                QB_TRY(make_y(Y));
                QB_TRY(qb_copy_matrix(ctx, st->m, kept, Y, st->rb, U, ldu, 0));
            }
            if (uinv) QB_TRY(qb_scale_rows_cols(ctx, U, U, st->m, kept, uinv, uinv_len, nullptr, 1));
        } else {  // U = P Xn
            if (ldu != st->m && uinv) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "svd: strided U with fused scaling");
            svd_emit_kernel<<<grid_cap(ctx, st->m * kept, 256), 256, 0, ctx->stream>>>(
                Xsrc, xld, st->m, kept, xperm, st->sigma_dev, xnorm, 0, 0, st->colperm_dev, U, ldu, uinv, uinv_len,
                1);
            QB_LAUNCH_CHECK(ctx);
        }
    }
    if (V) {
        if (!st->tall) {  // V = Y (n x kept): written as Vc = conj(Y) or Vh = Y^H
            c128* Y = ws.get<c128>((size_t)(st->rb * kept));
            if (!Y) QB_FAIL(ctx, QB200_E_CUDA, "svd: workspace allocation failed");
            QB_TRY(make_y(Y));
            if (vmode == 0) {
                if (vinv) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "svd: fused V scaling needs vmode 1");
                QB_TRY(qb_copy_matrix(ctx, st->n, kept, Y, st->rb, V, ldv, 2));
            } else {
                QB_TRY(qb_copy_matrix(ctx, st->n, kept, Y, st->rb, V, ldv, 1));
                if (vinv) {
                    if (ldv != kept) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "svd: strided Vh with fused scaling");
                    QB_TRY(qb_scale_rows_cols(ctx, V, V, kept, st->n, nullptr, 1, vinv, vinv_div));
                }
            }
        } else {  // V = P Xn
            svd_emit_kernel<<<grid_cap(ctx, st->n * kept, 256), 256, 0, ctx->stream>>>(
                Xsrc, xld, st->n, kept, xperm, st->sigma_dev, xnorm, 1, vmode, st->colperm_dev, V, ldv, vinv, 0,
                vinv ? vinv_div : 1);
            QB_LAUNCH_CHECK(ctx);
        }
    }
    if (S) {
        svd_emit_sigma_kernel<<<(unsigned)((kept + 255) / 256), 256, 0, ctx->stream>>>(st->sigma_dev, st->perm_dev, kept,
                                                                                      sigma_scale, S);
        QB_LAUNCH_CHECK(ctx);
    }
    return QB200_OK;
}

// ---------------------------------------------------------------------------------------------
// C-ABI: svd of the (left | right) matricisation of a tensor
extern "C" int32_t qb200_svd_last_sweeps(qb200_ctx* ctx) { return ctx ? ctx->last_svd_sweeps : -1; }
extern "C" int32_t qb200_svd_totals(qb200_ctx* ctx, int64_t* calls, int64_t* sweeps) {
    if (!ctx || !calls || !sweeps) return QB200_E_INVALID;
    *calls = ctx->svd_calls;
    *sweeps = ctx->svd_sweeps;
    for (qb200_ctx* w : ctx->workers) {
        *calls += w->svd_calls;
        *sweeps += w->svd_sweeps;
    }
    return QB200_OK;
}

int32_t qb_matricize(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* order, int32_t nleft, Workspace& ws,
                     const c128** mat, int64_t* rows, int64_t* cols) {
    if (!A || (A->dtype != QB200_C128 && A->dtype != QB200_C64))
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "factorisation: complex tensor required");
    if (nleft < 0 || nleft > A->rank) QB_FAIL(ctx, QB200_E_INVALID, "factorisation: bad nleft");
    bool seen[QB200_MAX_RANK] = {false}, identity = true;
    int64_t r = 1, c = 1;
    for (int i = 0; i < A->rank; ++i) {
        int p = order[i];
        if (p < 0 || p >= A->rank || seen[p]) QB_FAIL(ctx, QB200_E_INVALID, "factorisation: order is not a permutation");
        seen[p] = true;
        if (p != i) identity = false;
        (i < nleft ? r : c) *= A->ext[p];
    }
    *rows = r;
    *cols = c;
    if (identity && A->dtype == QB200_C128) {
        *mat = (const c128*)A->data;
        return QB200_OK;
    }
    c128* tmp = ws.get<c128>((size_t)A->numel());
    if (!tmp) QB_FAIL(ctx, QB200_E_CUDA, "factorisation: workspace allocation failed");
    *mat = tmp;
    if (A->dtype == QB200_C64) {
        // ComplexF32 tensors are factorised in FP64: permute in FP32 (half the bytes), then widen
        const void* src = A->data;
        if (!identity) {
            float2* p32 = ws.get<float2>((size_t)A->numel());
            if (!p32) QB_FAIL(ctx, QB200_E_CUDA, "factorisation: workspace allocation failed");
            qb200_tensor view = *A, outv = *A;
            outv.data = p32;
            for (int i = 0; i < A->rank; ++i) outv.ext[i] = A->ext[order[i]];
            QB_TRY(qb200_permute(ctx, &view, order, &outv));
            src = p32;
        }
        return qb_widen_c64(ctx, src, tmp, A->numel());
    }
    qb200_tensor view = *A;
    qb200_tensor outv = *A;
    outv.data = tmp;
    for (int i = 0; i < A->rank; ++i) outv.ext[i] = A->ext[order[i]];
    QB_TRY(qb200_permute(ctx, &view, order, &outv));
    return QB200_OK;
}

extern "C" int32_t qb200_svd(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* order, int32_t nleft,
                             int64_t maxdim, double threshold, qb200_tensor* U, qb200_tensor* S, qb200_tensor* Vc,
                             int64_t* kept_out, double* discarded_weight) {
    if (!ctx || !A || !order || !U || !S || !Vc) QB_FAIL(ctx, QB200_E_INVALID, "svd: null argument");
    Workspace ws(ctx);
    const c128* mat;
    int64_t m, n;
    QB_TRY(qb_matricize(ctx, A, order, nleft, ws, &mat, &m, &n));
    int64_t k = std::min(m, n);
    const bool c64 = (A->dtype == QB200_C64);
    if (U->dtype != A->dtype || Vc->dtype != A->dtype || S->dtype != QB200_F64)
        QB_FAIL(ctx, QB200_E_INVALID, "svd: U, Vc must have the type of A and S must be real");
    if (U->numel() < m * k || Vc->numel() < n * k || S->numel() < k)
        QB_FAIL(ctx, QB200_E_INVALID, "svd: outputs too small (need k = min(rows, cols) columns)");
    SvdState* st = nullptr;
    std::vector<double> sigma;
    QB_TRY(qb_svd_factor(ctx, m, n, mat, m, &st, sigma));
    // truncate! rule (Chain.jl:404-417): i <= min(k, maxdim) and s[i] > threshold
    int64_t lim = (maxdim > 0) ? std::min(k, maxdim) : k;
    int64_t kept = 0;
    for (int64_t i = 0; i < lim; ++i)
        if (threshold < 0.0 || std::fabs(sigma[i]) > threshold) kept++;
        else break;
    double dw = 0.0;
    for (int64_t i = k - 1; i >= kept; --i) dw += sigma[i] * sigma[i];
    c128 *Uw = (c128*)U->data, *Vw = (c128*)Vc->data;
    if (c64) {  // factors leave the FP64 factorisation through a workspace and are narrowed to float2
        Uw = ws.get<c128>((size_t)std::max<int64_t>(m * kept, 1));
        Vw = ws.get<c128>((size_t)std::max<int64_t>(n * kept, 1));
        if (!Uw || !Vw) {
            qb_svd_release(ctx, st);
            QB_FAIL(ctx, QB200_E_CUDA, "svd: workspace allocation failed");
        }
    }
    int32_t r = qb_svd_emit(ctx, st, kept, Uw, m, (double*)S->data, Vw, n, 0, nullptr, 0, nullptr, 1, 1.0);
    qb_svd_release(ctx, st);
    QB_TRY(r);
    if (c64) {
        QB_TRY(qb_narrow_c128(ctx, Uw, U->data, m * kept));
        QB_TRY(qb_narrow_c128(ctx, Vw, Vc->data, n * kept));
    }
    // shrink the trailing (bond) extent of the outputs to `kept`
    U->ext[U->rank - 1] = kept;
    Vc->ext[Vc->rank - 1] = kept;
    S->ext[S->rank - 1] = kept;
    if (kept_out) *kept_out = kept;
    if (discarded_weight) *discarded_weight = dw;
    return QB200_OK;
}


constexpr size_t CHOL_SMEM = (size_t)2 * JP * GLD * sizeof(c128);

// One Cholesky-QR step on the 64-column panel P (m x 64, ld, m % 64 == 0): P <- P R^-1, R (64 x 64, ldr) written.
// Raises *fail_dev when the panel is too ill conditioned for a Gram-based step (caller falls back to TSQR).
int32_t qb_cholqr_panel_step(qb200_ctx* ctx, c128* P, int64_t ld, int64_t m, c128* R, int64_t ldr, c128* Gpart,
                             c128* Wbuf, int* flags_dev, int* fail_dev, const double* before2_dev, double dep_tol) {
    const int mp = (int)m, nchunk = mp / G_BKR;
    // few, fat CTAs: the single-CTA Cholesky kernel has to sum every partial Gram
    const int gram_ctas = std::max(1, std::min(16, nchunk));
    jacobi_gram_kernel<0><<<gram_ctas, 256, GRAM_SMEM, ctx->stream>>>(P, ld, mp, 2, -1, 1, Gpart);
    QB_LAUNCH_CHECK(ctx);
    panel_chol_kernel<<<1, 256, CHOL_SMEM, ctx->stream>>>(Gpart, gram_ctas, nchunk, Wbuf, R, ldr, flags_dev, fail_dev, 1e-11,
                                                          before2_dev, dep_tol * dep_tol);
    QB_LAUNCH_CHECK(ctx);
    const int u_nchunk = mp / U_ROWS;
    const int upd_ctas = std::max(1, std::min(2 * ctx->sm_count, u_nchunk));
    (update_3m() ? jacobi_update_kernel<1> : jacobi_update_kernel<0>)<<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(P, ld, 2, -1, Wbuf, flags_dev, 1, u_nchunk, nullptr);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}
size_t qb_cholqr_gpart_elems(qb200_ctx* ctx) { return (size_t)2 * 2 * ctx->sm_count * JP * JP; }

// Timing harness for the diagnostics library (qb200_bench_update_variants): `steps` launches of the Jacobi update kernel
// on a k x k matrix with W = I for every pair, per variant (see the VAR template parameter); ms_out[v] = mean
// microseconds per launch.  Variant 0 is the production kernel.
namespace qb {
int32_t qb_update_bench(qb200_ctx* ctx, int k, int steps, double* us_out, int nvar) {
    const int np = (k + 63) / 64 * 64, nb = np / JB, npairs = nb / 2;
    Workspace ws(ctx);
    c128* Z = ws.get<c128>((size_t)np * np);
    c128* Wg = ws.get<c128>((size_t)npairs * JP * JP);
    int* flags = ws.get<int>(npairs);
    if (!Z || !Wg || !flags) QB_FAIL(ctx, QB200_E_CUDA, "update bench: allocation failed");
    svd_init_kernel<<<grid_cap(ctx, (int64_t)np * np, 256), 256, 0, ctx->stream>>>(Z, np, np, np, 0, 0, nullptr, 1, 0);
    std::vector<c128> w((size_t)npairs * JP * JP, make_double2(0.0, 0.0));
    for (int p = 0; p < npairs; ++p)
        for (int i = 0; i < JP; ++i) w[(size_t)p * JP * JP + i * JP + i] = make_double2(1.0, 0.0);
    std::vector<int> f(npairs, 1);
    QB_CUDA(ctx, cudaMemcpyAsync(Wg, w.data(), sizeof(c128) * w.size(), cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, cudaMemcpyAsync(flags, f.data(), sizeof(int) * npairs, cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int u_nchunk = np / U_ROWS;
    const int upd_ctas = std::max(1, std::min(2 * ctx->sm_count, npairs * u_nchunk));
    auto launch = [&](int v, int step) {
        switch (v) {
            case 0: jacobi_update_kernel<1, 0><<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(Z, np, nb, step, Wg, flags, npairs, u_nchunk, nullptr); break;
            case 1: jacobi_update_kernel<1, 1><<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(Z, np, nb, step, Wg, flags, npairs, u_nchunk, nullptr); break;
            case 2: jacobi_update_kernel<1, 2><<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(Z, np, nb, step, Wg, flags, npairs, u_nchunk, nullptr); break;
            case 3: jacobi_update_kernel<1, 4><<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(Z, np, nb, step, Wg, flags, npairs, u_nchunk, nullptr); break;
            case 4: jacobi_update_kernel<1, 7><<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(Z, np, nb, step, Wg, flags, npairs, u_nchunk, nullptr); break;
            case 5: jacobi_update_kernel<1, 15><<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(Z, np, nb, step, Wg, flags, npairs, u_nchunk, nullptr); break;
            default: jacobi_update_kernel<0, 0><<<upd_ctas, UPD_THREADS, UPD_SMEM, ctx->stream>>>(Z, np, nb, step, Wg, flags, npairs, u_nchunk, nullptr); break;
        }
    };
    for (int v = 0; v < nvar && v < 7; ++v) {
        if (v >= 1 && v <= 5) {
            cudaError_t e = cudaSuccess;
            if (v == 1) e = cudaFuncSetAttribute(jacobi_update_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
            if (v == 2) e = cudaFuncSetAttribute(jacobi_update_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
            if (v == 3) e = cudaFuncSetAttribute(jacobi_update_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
            if (v == 4) e = cudaFuncSetAttribute(jacobi_update_kernel<1, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
            if (v == 5) e = cudaFuncSetAttribute(jacobi_update_kernel<1, 15>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UPD_SMEM);
            QB_CUDA(ctx, e);
        }
        for (int i = 0; i < 5; ++i) launch(v, i % (nb - 1));
        QB_LAUNCH_CHECK(ctx);
        QB_TRY(qb200_timer_begin(ctx));
        for (int i = 0; i < steps; ++i) launch(v, i % (nb - 1));
        QB_LAUNCH_CHECK(ctx);
        double ms = 0.0;
        QB_TRY(qb200_timer_end(ctx, &ms));
        us_out[v] = ms * 1e3 / steps;
    }
    return QB200_OK;
}
}  // namespace qb
