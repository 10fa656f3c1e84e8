// K2/K3/K6/K7/K8: HBM-bound helpers -- mode scale (with pinv), slice, select, conj, permute, norms,
// gate application on the two-site wave function.  All are single-pass, 16-byte vectorised
// (one ComplexF64 per access), grid-stride with grids sized as multiples of the SM count.
#include <algorithm>

#include "common.cuh"
#include "mma.cuh"

using namespace qb;

static inline unsigned grid_for(qb200_ctx* ctx, int64_t n, int threads) {
    int64_t blocks = (n + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

// out[i, j, o] = in[i, j, o] * f(vec[j])
template <typename T>  // T = double2 (ComplexF64) or float2 (ComplexF32); the factor is applied in FP64
__global__ void scale_mode_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t inner, int64_t d,
                                  int64_t total, const double* __restrict__ vec, int inverse, double atol) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t j = (idx / inner) % d;
        double v = vec[j];
        if (inverse) v = (fabs(v) > atol) ? 1.0 / v : 0.0;
        T x = in[idx];
        T o;
        o.x = (decltype(o.x))(x.x * v);
        o.y = (decltype(o.y))(x.y * v);
        out[idx] = o;
    }
}

// ComplexF32, two elements (16 bytes) per thread; IDX = uint32_t when the tensor has < 2^31 elements (32-bit
// divisions instead of emulated 64-bit ones: the kernel is index-arithmetic bound otherwise)
template <typename IDX>
__global__ void scale_mode_c64x2_kernel(const float4* __restrict__ in, float4* __restrict__ out, IDX inner, IDX d,
                                        IDX npairs, const double* __restrict__ vec, int inverse, double atol) {
    for (IDX idx = (IDX)blockIdx.x * blockDim.x + threadIdx.x; idx < npairs; idx += (IDX)gridDim.x * blockDim.x) {
        const IDX e = 2 * idx;
        double v0 = vec[(e / inner) % d], v1 = vec[((e + 1) / inner) % d];
        if (inverse) {
            v0 = (fabs(v0) > atol) ? 1.0 / v0 : 0.0;
            v1 = (fabs(v1) > atol) ? 1.0 / v1 : 0.0;
        }
        float4 x = in[idx];
        out[idx] = make_float4((float)(x.x * v0), (float)(x.y * v0), (float)(x.z * v1), (float)(x.w * v1));
    }
}

int32_t qb_scale_mode_raw(qb200_ctx* ctx, const c128* in, c128* out, int64_t inner, int64_t d, int64_t outer,
                          const double* vec, int inverse, double atol) {
    int64_t total = inner * d * outer;
    if (total == 0) return QB200_OK;
    scale_mode_kernel<c128><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(in, out, inner, d, total, vec, inverse, atol);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// out(r, c) = in(r, c) * rvec[r % rlen] * cvec[c / cdiv]   (either vector may be null)
__global__ void scale_rows_cols_kernel(const c128* __restrict__ in, c128* __restrict__ out, int64_t rows,
                                       int64_t total, const double* __restrict__ rvec, int64_t rlen,
                                       const double* __restrict__ cvec, int64_t cdiv) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = idx % rows, c = idx / rows;
        double v = 1.0;
        if (rvec) v *= rvec[r % rlen];
        if (cvec) v *= cvec[c / cdiv];
        c128 x = in[idx];
        out[idx] = make_double2(x.x * v, x.y * v);
    }
}

int32_t qb_scale_rows_cols(qb200_ctx* ctx, const c128* in, c128* out, int64_t rows, int64_t cols,
                           const double* rvec, int64_t rlen, const double* cvec, int64_t cdiv) {
    int64_t total = rows * cols;
    if (total == 0) return QB200_OK;
    scale_rows_cols_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(in, out, rows, total, rvec, rlen, cvec,
                                                                             cdiv);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// B = A (ct = 0), B = A^H (ct = 1, B is n x m) or B = conj(A) (ct = 2); tiled through shared memory
__global__ void copy_matrix_kernel(int64_t m, int64_t n, const c128* __restrict__ A, int64_t lda,
                                   c128* __restrict__ B, int64_t ldb, int ct) {
    __shared__ c128 tile[32][33];
    int64_t i0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
        int64_t i = i0 + threadIdx.x, j = j0 + jj;
        if (i < m && j < n) tile[jj][threadIdx.x] = A[i + j * lda];
    }
    __syncthreads();
    if (ct != 1) {  // ct == 2: conjugate without transposing
        for (int jj = threadIdx.y; jj < 32; jj += blockDim.y) {
            int64_t i = i0 + threadIdx.x, j = j0 + jj;
            if (i < m && j < n) B[i + j * ldb] = (ct == 2) ? cconj(tile[jj][threadIdx.x]) : tile[jj][threadIdx.x];
        }
    } else {
        for (int ii = threadIdx.y; ii < 32; ii += blockDim.y) {
            int64_t j = j0 + threadIdx.x, i = i0 + ii;  // B(j, i) = conj(A(i, j))
            if (i < m && j < n) B[j + i * ldb] = cconj(tile[threadIdx.x][ii]);
        }
    }
}

int32_t qb_copy_matrix(qb200_ctx* ctx, int64_t m, int64_t n, const c128* A, int64_t lda, c128* B, int64_t ldb,
                       int ct) {
    if (m == 0 || n == 0) return QB200_OK;
    dim3 grid((unsigned)((m + 31) / 32), (unsigned)((n + 31) / 32));
    if (grid.y > 65535) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "copy_matrix: too many columns");
    copy_matrix_kernel<<<grid, dim3(32, 8), 0, ctx->stream>>>(m, n, A, lda, B, ldb, ct);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// θ[(l,o1),(o2,r)] <- sum_{i1,i2} G[o1,o2,i1,i2] θ[(l,i1),(i2,r)]  in place; G column-major (o1,o2,i1,i2).
// θ is (2χl) x (2χr) column-major, l fastest within rows, o2 fastest within columns.
__global__ void apply_gate2_kernel(c128* __restrict__ theta, int64_t chil, int64_t chir, const c128* __restrict__ G) {
    __shared__ c128 g[16];
    if (threadIdx.x < 16) g[threadIdx.x] = G[threadIdx.x];
    __syncthreads();
    int64_t total = chil * chir, ld = 2 * chil;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t l = idx % chil, r = idx / chil;
        c128* base = theta + l + (2 * r) * ld;
        c128 x[4];  // index i1 + 2*i2
        x[0] = base[0];
        x[1] = base[chil];
        x[2] = base[ld];
        x[3] = base[ld + chil];
        c128 y[4];
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            c128 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc = cadd(acc, cmul(g[o + 4 * i], x[i]));
            y[o] = acc;
        }
        base[0] = y[0];
        base[chil] = y[1];
        base[ld] = y[2];
        base[ld + chil] = y[3];
    }
}

int32_t qb_apply_gate2(qb200_ctx* ctx, c128* theta, int64_t chil, int64_t chir, const c128* gate_dev) {
    int64_t total = chil * chir;
    if (total == 0) return QB200_OK;
    apply_gate2_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(theta, chil, chir, gate_dev);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// t[i, o, r] <- sum_j G[o, j] t[i, j, r], in place, p <= 8
__global__ void apply_gate1_kernel(c128* __restrict__ t, int64_t inner, int p, int64_t outer, const c128* __restrict__ G) {
    __shared__ c128 g[64];
    if (threadIdx.x < p * p) g[threadIdx.x] = G[threadIdx.x];
    __syncthreads();
    int64_t total = inner * outer;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = idx % inner, r = idx / inner;
        c128* base = t + i + r * inner * p;
        c128 x[8], y[8];
        for (int j = 0; j < p; ++j) x[j] = base[j * inner];
        for (int o = 0; o < p; ++o) {
            c128 acc = make_double2(0.0, 0.0);
            for (int j = 0; j < p; ++j) acc = cadd(acc, cmul(g[o + p * j], x[j]));
            y[o] = acc;
        }
        for (int o = 0; o < p; ++o) base[o * inner] = y[o];
    }
}

int32_t qb_apply_gate1(qb200_ctx* ctx, c128* t, int64_t inner, int64_t p, int64_t outer, const c128* gate_dev) {
    if (p > 8) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "physical dimension %lld > 8", (long long)p);
    int64_t total = inner * outer;
    if (total == 0) return QB200_OK;
    apply_gate1_kernel<<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(t, inner, (int)p, outer, gate_dev);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// deterministic two-pass sum of squares
__global__ void sumsq_partial_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ partial) {
    __shared__ double sh[32];
    double acc = 0.0;
    if (n % 2 == 0 && ((uintptr_t)x & 15) == 0) {  // 16-byte loads, two running sums
        const double2* x2 = reinterpret_cast<const double2*>(x);
        double acc2 = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 2;
             i += (int64_t)gridDim.x * blockDim.x) {
            double2 v = x2[i];
            acc += v.x * v.x;
            acc2 += v.y * v.y;
        }
        acc += acc2;
    } else
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = x[i];
        acc += v * v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}
__global__ void sum_final_kernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partial[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) out[0] = v;
    }
}

int32_t qb_sumsq(qb200_ctx* ctx, const double* x, int64_t n, double* result_host) {
    Workspace ws(ctx);
    unsigned blocks = grid_for(ctx, n, 256);
    double* partial = ws.get<double>(blocks + 1);
    if (!partial) QB_FAIL(ctx, QB200_E_CUDA, "workspace allocation failed");
    sumsq_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(x, n, partial);
    QB_LAUNCH_CHECK(ctx);
    sum_final_kernel<<<1, 256, 0, ctx->stream>>>(partial, (int)blocks, partial + blocks);
    QB_LAUNCH_CHECK(ctx);
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, partial + blocks, sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    *result_host = ctx->scratch_host[0];
    return QB200_OK;
}

// sum of squares of FP32 data with FP64 accumulation (norm of a ComplexF32 tensor), same two-pass scheme
__global__ void sumsq_f32_partial_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ partial) {
    double s = 0.0;
    if (n % 4 == 0 && ((uintptr_t)x & 15) == 0) {  // 16-byte loads
        const float4* x4 = reinterpret_cast<const float4*>(x);
        double s2 = 0.0;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n / 4;
             i += (int64_t)gridDim.x * blockDim.x) {
            float4 v = x4[i];
            s += (double)v.x * v.x + (double)v.y * v.y;
            s2 += (double)v.z * v.z + (double)v.w * v.w;
        }
        s += s2;
    } else
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = (double)x[i];
        s += v * v;
    }
    s = warp_sum(s);
    __shared__ double sh[32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = s;
    __syncthreads();
    if (w == 0) {
        s = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0.0;
        s = warp_sum(s);
        if (lane == 0) partial[blockIdx.x] = s;
    }
}

static int32_t qb_sumsq_f32(qb200_ctx* ctx, const float* x, int64_t n, double* result_host) {
    *result_host = 0.0;
    if (n == 0) return QB200_OK;
    Workspace ws(ctx);
    int blocks = (int)grid_for(ctx, n, 256);
    double* partial = ws.get<double>((size_t)blocks + 1);
    if (!partial) QB_FAIL(ctx, QB200_E_CUDA, "norm: workspace allocation failed");
    sumsq_f32_partial_kernel<<<blocks, 256, 0, ctx->stream>>>(x, n, partial);
    QB_LAUNCH_CHECK(ctx);
    sum_final_kernel<<<1, 256, 0, ctx->stream>>>(partial, blocks, partial + blocks);
    QB_LAUNCH_CHECK(ctx);
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, partial + blocks, sizeof(double), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    QB_CUDA(ctx, qb_stream_sync(ctx));
    *result_host = ctx->scratch_host[0];
    return QB200_OK;
}

__global__ void widen_c64_kernel(const float2* __restrict__ src, c128* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float2 v = src[i];
        dst[i] = make_double2((double)v.x, (double)v.y);
    }
}
__global__ void narrow_c128_kernel(const c128* __restrict__ src, float2* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        c128 v = src[i];
        dst[i] = make_float2((float)v.x, (float)v.y);
    }
}
int32_t qb_widen_c64(qb200_ctx* ctx, const void* src, c128* dst, int64_t n) {
    if (n <= 0) return QB200_OK;
    widen_c64_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>((const float2*)src, dst, n);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}
int32_t qb_narrow_c128(qb200_ctx* ctx, const c128* src, void* dst, int64_t n) {
    if (n <= 0) return QB200_OK;
    narrow_c128_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>(src, (float2*)dst, n);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

// generic gather: out[idx] = in[sum_j coord_j(idx) * stride_j] (optionally conjugated); out is dense
struct GatherModes {
    int n;
    int64_t ext[QB200_MAX_RANK];
    int64_t stride[QB200_MAX_RANK];
};
template <typename T, bool CPLX>
__global__ void gather_kernel(const T* __restrict__ in, T* __restrict__ out, const GatherModes gm, int64_t total,
                              int64_t base, int conj) {
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = idx, off = base;
        for (int j = 0; j < gm.n; ++j) {
            int64_t c = rem % gm.ext[j];
            rem /= gm.ext[j];
            off += c * gm.stride[j];
        }
        T v = in[off];
        if constexpr (CPLX) {
            if (conj) v.y = -v.y;
        }
        out[idx] = v;
    }
}

// Fast path of the gather (every slice / select / conj / permute that fits): modes that are contiguous after one
// another in the input are merged on the host, ComplexF32 / Float64 data whose fastest output mode is unit-stride in
// the input moves as 16-byte pairs, indices are 32-bit and decomposed with multiply-high "magic" divisions
// (q = umulhi(n, mul) >> shr, exact for n < 2^31) instead of 64-bit div/mod chains, and every thread keeps four
// independent 16-byte loads in flight.  KIND: 0 real, 1 one complex number per element, 2 two complex numbers (float4).
constexpr int FG_MAX = 16;
struct FastGather {
    int n;
    uint32_t ext[FG_MAX], mul[FG_MAX], shr[FG_MAX];
    int64_t stride[FG_MAX];
};
template <typename T, int KIND>
__global__ void __launch_bounds__(256) gather_fast_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                          const FastGather fg, uint32_t total, int64_t base, int conj) {
    constexpr int ILP = 4;
    for (uint32_t idx0 = blockIdx.x * (256u * ILP) + threadIdx.x; idx0 < total; idx0 += gridDim.x * (256u * ILP)) {
        T v[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const uint32_t idx = idx0 + 256u * u;
            if (idx < total) {
                uint32_t rem = idx;
                int64_t off = base;
                for (int j = 0; j < fg.n - 1; ++j) {
                    const uint32_t q = fg.ext[j] == 1 ? rem : (__umulhi(rem, fg.mul[j]) >> fg.shr[j]);
                    off += (int64_t)(rem - q * fg.ext[j]) * fg.stride[j];
                    rem = q;
                }
                off += (int64_t)rem * fg.stride[fg.n - 1];
                v[u] = in[off];
            }
        }
#pragma unroll
        for (int u = 0; u < ILP; ++u) {
            const uint32_t idx = idx0 + 256u * u;
            if (idx < total) {
                if constexpr (KIND >= 1) {
                    if (conj) v[u].y = -v[u].y;
                }
                if constexpr (KIND == 2) {
                    if (conj) v[u].w = -v[u].w;
                }
                out[idx] = v[u];
            }
        }
    }
}

// Transposing gathers (the fastest output mode is strided in the input while another output mode is unit-stride in
// the input, e.g. permutedims (l,o,r) -> (r,o,l)): 32 x 32 tiles through shared memory so that both the reads (along
// the input-contiguous mode) and the writes (along the output-contiguous mode) are full 512-byte / 256-byte warp
// transactions instead of one element per 32-byte sector.  The remaining modes are enumerated per tile.
struct TileGather {
    int n;  // modes other than the two tiled ones
    uint32_t ext[FG_MAX], mul[FG_MAX], shr[FG_MAX];
    int64_t istride[FG_MAX], ostride[FG_MAX];
    uint32_t e0, ec, tiles0, tilesc, mul0, shr0, mulc, shrc;
    int64_t s0;  // input stride of output mode 0 (its output stride is 1)
    int64_t oc;  // output stride of the input-contiguous mode (its input stride is 1)
};
template <typename T, int KIND>
__global__ void __launch_bounds__(256) gather_tiled_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                           const TileGather tg, uint32_t ntiles, int64_t base, int conj) {
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (uint32_t tid = blockIdx.x; tid < ntiles; tid += gridDim.x) {
        uint32_t q = tg.tiles0 == 1 ? tid : (__umulhi(tid, tg.mul0) >> tg.shr0);
        const uint32_t t0 = tid - q * tg.tiles0;
        uint32_t rest = tg.tilesc == 1 ? q : (__umulhi(q, tg.mulc) >> tg.shrc);
        const uint32_t tc = q - rest * tg.tilesc;
        int64_t ioff = base, ooff = 0;
        for (int j = 0; j < tg.n; ++j) {
            const uint32_t qq = tg.ext[j] == 1 ? rest : (__umulhi(rest, tg.mul[j]) >> tg.shr[j]);
            const uint32_t c = rest - qq * tg.ext[j];
            ioff += (int64_t)c * tg.istride[j];
            ooff += (int64_t)c * tg.ostride[j];
            rest = qq;
        }
        {
            const uint32_t ic = tc * 32 + tx;  // along the input-contiguous mode
#pragma unroll
            for (int r = 0; r < 32; r += 8) {
                const uint32_t i0 = t0 * 32 + ty + r;
                if (i0 < tg.e0 && ic < tg.ec) tile[ty + r][tx] = in[ioff + (int64_t)i0 * tg.s0 + ic];
            }
        }
        __syncthreads();
        {
            const uint32_t i0 = t0 * 32 + tx;  // along the output-contiguous mode
#pragma unroll
            for (int r = 0; r < 32; r += 8) {
                const uint32_t ic = tc * 32 + ty + r;
                if (i0 < tg.e0 && ic < tg.ec) {
                    T v = tile[tx][ty + r];
                    if constexpr (KIND >= 1) {
                        if (conj) v.y = -v.y;
                    }
                    out[ooff + i0 + (int64_t)ic * tg.oc] = v;
                }
            }
        }
        __syncthreads();
    }
}

static void make_fastdiv(uint32_t d, uint32_t* mul, uint32_t* shr) {
    if (d <= 1) {
        *mul = 0;
        *shr = 0;
        return;
    }
    int l = 0;
    while ((1ull << l) < d) ++l;  // ceil(log2 d)
    const int p = 31 + l;
    *mul = (uint32_t)(((1ull << p) + d - 1) / d);
    *shr = (uint32_t)(p - 32);
}

template <typename T, int KIND>
static void launch_gather_fast(qb200_ctx* ctx, const void* in, void* out, const FastGather& fg, int64_t total, int64_t base,
                               int conj) {
    int64_t blocks = (total + 1023) / 1024, cap = (int64_t)ctx->sm_count * 8;
    gather_fast_kernel<T, KIND><<<(unsigned)std::max<int64_t>(1, std::min(blocks, cap)), 256, 0, ctx->stream>>>(
        (const T*)in, (T*)out, fg, (uint32_t)total, base, conj);
}

static int32_t gather(qb200_ctx* ctx, const qb200_tensor* A, qb200_tensor* out, const GatherModes& gm_in, int64_t base,
                      int conj) {
    int64_t total = 1;
    for (int j = 0; j < gm_in.n; ++j) total *= gm_in.ext[j];
    if (total == 0) return QB200_OK;
    // merge: drop extent-1 modes, fuse a mode into its predecessor when it continues it contiguously in the input
    GatherModes gm;
    gm.n = 0;
    for (int j = 0; j < gm_in.n; ++j) {
        if (gm_in.ext[j] == 1) continue;
        if (gm.n > 0 && gm_in.stride[j] == gm.stride[gm.n - 1] * gm.ext[gm.n - 1]) {
            gm.ext[gm.n - 1] *= gm_in.ext[j];
            continue;
        }
        gm.ext[gm.n] = gm_in.ext[j];
        gm.stride[gm.n] = gm_in.stride[j];
        gm.n++;
    }
    if (gm.n == 0) {
        gm.n = 1;
        gm.ext[0] = 1;
        gm.stride[0] = 1;
    }
    // transposing gather: tile through shared memory
    if (gm.n >= 2 && gm.stride[0] != 1 && gm.ext[0] >= 8) {
        int jc = -1;
        for (int j = 1; j < gm.n; ++j)
            if (gm.stride[j] == 1 && gm.ext[j] >= 8) jc = j;
        const int64_t tiles0 = (gm.ext[0] + 31) / 32, tilesc = jc > 0 ? (gm.ext[jc] + 31) / 32 : 0;
        int64_t ntiles = tiles0 * tilesc;
        for (int j = 1; j < gm.n && jc > 0; ++j)
            if (j != jc) ntiles *= gm.ext[j];
        if (jc > 0 && gm.n - 2 <= FG_MAX && ntiles < (1ll << 31) && gm.ext[0] < (1ll << 31) && gm.ext[jc] < (1ll << 31)) {
            TileGather tg;
            tg.n = 0;
            int64_t ostr = 1;
            for (int j = 0; j < gm.n; ++j) {
                if (j == jc) tg.oc = ostr;
                if (j != 0 && j != jc) {
                    tg.ext[tg.n] = (uint32_t)gm.ext[j];
                    tg.istride[tg.n] = gm.stride[j];
                    tg.ostride[tg.n] = ostr;
                    make_fastdiv((uint32_t)gm.ext[j], &tg.mul[tg.n], &tg.shr[tg.n]);
                    tg.n++;
                }
                ostr *= gm.ext[j];
            }
            tg.e0 = (uint32_t)gm.ext[0];
            tg.ec = (uint32_t)gm.ext[jc];
            tg.s0 = gm.stride[0];
            tg.tiles0 = (uint32_t)tiles0;
            tg.tilesc = (uint32_t)tilesc;
            make_fastdiv(tg.tiles0, &tg.mul0, &tg.shr0);
            make_fastdiv(tg.tilesc, &tg.mulc, &tg.shrc);
            const unsigned blocks = (unsigned)std::min<int64_t>(ntiles, (int64_t)ctx->sm_count * 16);
            if (A->dtype == QB200_C128)
                gather_tiled_kernel<c128, 1><<<blocks, 256, 0, ctx->stream>>>((const c128*)A->data, (c128*)out->data, tg,
                                                                              (uint32_t)ntiles, base, conj);
            else if (A->dtype == QB200_C64)
                gather_tiled_kernel<float2, 1><<<blocks, 256, 0, ctx->stream>>>(
                    (const float2*)A->data, (float2*)out->data, tg, (uint32_t)ntiles, base, conj);
            else
                gather_tiled_kernel<double, 0><<<blocks, 256, 0, ctx->stream>>>(
                    (const double*)A->data, (double*)out->data, tg, (uint32_t)ntiles, base, 0);
            QB_LAUNCH_CHECK(ctx);
            return QB200_OK;
        }
    }
    // 16-byte pairs for 8-byte element types
    bool pairs = false;
    if (A->dtype != QB200_C128 && gm.stride[0] == 1 && gm.ext[0] % 2 == 0 && base % 2 == 0 &&
        ((uintptr_t)A->data % 16 == 0) && ((uintptr_t)out->data % 16 == 0)) {
        pairs = true;
        for (int j = 1; j < gm.n; ++j) pairs = pairs && (gm.stride[j] % 2 == 0);
    }
    int64_t vtotal = pairs ? total / 2 : total;
    if (gm.n <= FG_MAX && vtotal < (1ll << 31)) {
        FastGather fg;
        fg.n = gm.n;
        for (int j = 0; j < gm.n; ++j) {
            int64_t e = gm.ext[j], st = gm.stride[j];
            if (pairs) {
                if (j == 0) e /= 2;
                else st /= 2;
            }
            fg.ext[j] = (uint32_t)e;
            fg.stride[j] = st;
            make_fastdiv((uint32_t)e, &fg.mul[j], &fg.shr[j]);
        }
        const int64_t vbase = pairs ? base / 2 : base;
        if (A->dtype == QB200_C128)
            launch_gather_fast<c128, 1>(ctx, A->data, out->data, fg, vtotal, vbase, conj);
        else if (A->dtype == QB200_C64 && pairs)
            launch_gather_fast<float4, 2>(ctx, A->data, out->data, fg, vtotal, vbase, conj);
        else if (A->dtype == QB200_C64)
            launch_gather_fast<float2, 1>(ctx, A->data, out->data, fg, vtotal, vbase, conj);
        else if (pairs)
            launch_gather_fast<double2, 0>(ctx, A->data, out->data, fg, vtotal, vbase, 0);
        else
            launch_gather_fast<double, 0>(ctx, A->data, out->data, fg, vtotal, vbase, 0);
        QB_LAUNCH_CHECK(ctx);
        return QB200_OK;
    }
    // general fallback (more than 2^31 elements or more than FG_MAX non-mergeable modes): 64-bit index arithmetic
    if (A->dtype == QB200_C128)
        gather_kernel<c128, true><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(
            (const c128*)A->data, (c128*)out->data, gm, total, base, conj);
    else if (A->dtype == QB200_C64)
        gather_kernel<float2, true><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(
            (const float2*)A->data, (float2*)out->data, gm, total, base, conj);
    else
        gather_kernel<double, false><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(
            (const double*)A->data, (double*)out->data, gm, total, base, 0);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

static void dense_strides(const qb200_tensor* A, int64_t* st) {
    int64_t s = 1;
    for (int i = 0; i < A->rank; ++i) {
        st[i] = s;
        s *= A->ext[i];
    }
}

extern "C" {

int32_t qb200_scale_mode(qb200_ctx* ctx, const qb200_tensor* A, int32_t mode_pos, const qb200_tensor* vec,
                         int32_t inverse, double atol, qb200_tensor* out) {
    if (!A || !vec || !out || mode_pos < 0 || mode_pos >= A->rank) QB_FAIL(ctx, QB200_E_INVALID, "scale_mode: bad argument");
    if ((A->dtype != QB200_C128 && A->dtype != QB200_C64) || vec->dtype != QB200_F64 || out->dtype != A->dtype)
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "scale_mode: needs a complex tensor (same type in and out) and a real vector");
    if (vec->numel() != A->ext[mode_pos] || out->numel() != A->numel())
        QB_FAIL(ctx, QB200_E_INVALID, "scale_mode: extent mismatch");
    int64_t inner = 1, outer = 1;
    for (int i = 0; i < mode_pos; ++i) inner *= A->ext[i];
    for (int i = mode_pos + 1; i < A->rank; ++i) outer *= A->ext[i];
    if (A->dtype == QB200_C64) {
        int64_t total = inner * A->ext[mode_pos] * outer;
        if (total == 0) return QB200_OK;
        if (total % 2 == 0 && (uintptr_t)A->data % 16 == 0 && (uintptr_t)out->data % 16 == 0) {
            if (total < (1ll << 31))
                scale_mode_c64x2_kernel<uint32_t><<<grid_for(ctx, total / 2, 256), 256, 0, ctx->stream>>>(
                    (const float4*)A->data, (float4*)out->data, (uint32_t)inner, (uint32_t)A->ext[mode_pos],
                    (uint32_t)(total / 2), (const double*)vec->data, inverse, atol);
            else
                scale_mode_c64x2_kernel<int64_t><<<grid_for(ctx, total / 2, 256), 256, 0, ctx->stream>>>(
                    (const float4*)A->data, (float4*)out->data, inner, A->ext[mode_pos], total / 2,
                    (const double*)vec->data, inverse, atol);
            QB_LAUNCH_CHECK(ctx);
            return QB200_OK;
        }
        scale_mode_kernel<float2><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(
            (const float2*)A->data, (float2*)out->data, inner, A->ext[mode_pos], total, (const double*)vec->data, inverse,
            atol);
        QB_LAUNCH_CHECK(ctx);
        return QB200_OK;
    }
    return qb_scale_mode_raw(ctx, (const c128*)A->data, (c128*)out->data, inner, A->ext[mode_pos], outer,
                             (const double*)vec->data, inverse, atol);
}

int32_t qb200_slice_mode(qb200_ctx* ctx, const qb200_tensor* A, int32_t mode_pos, int64_t count, qb200_tensor* out) {
    if (!A || !out || mode_pos < 0 || mode_pos >= A->rank || count < 0 || count > A->ext[mode_pos])
        QB_FAIL(ctx, QB200_E_INVALID, "slice_mode: bad argument");
    if (out->dtype != A->dtype) QB_FAIL(ctx, QB200_E_INVALID, "slice_mode: dtype mismatch");
    GatherModes gm;
    gm.n = A->rank;
    int64_t st[QB200_MAX_RANK];
    dense_strides(A, st);
    int64_t total = 1;
    for (int i = 0; i < A->rank; ++i) {
        gm.ext[i] = (i == mode_pos) ? count : A->ext[i];
        gm.stride[i] = st[i];
        total *= gm.ext[i];
    }
    if (out->numel() != total) QB_FAIL(ctx, QB200_E_INVALID, "slice_mode: output has wrong size");
    return gather(ctx, A, out, gm, 0, 0);
}

int32_t qb200_select_mode(qb200_ctx* ctx, const qb200_tensor* A, int32_t mode_pos, int64_t index, qb200_tensor* out) {
    if (!A || !out || mode_pos < 0 || mode_pos >= A->rank || index < 0 || index >= A->ext[mode_pos])
        QB_FAIL(ctx, QB200_E_INVALID, "select_mode: bad argument");
    if (out->dtype != A->dtype) QB_FAIL(ctx, QB200_E_INVALID, "select_mode: dtype mismatch");
    GatherModes gm;
    gm.n = 0;
    int64_t st[QB200_MAX_RANK];
    dense_strides(A, st);
    int64_t total = 1;
    for (int i = 0; i < A->rank; ++i) {
        if (i == mode_pos) continue;
        gm.ext[gm.n] = A->ext[i];
        gm.stride[gm.n] = st[i];
        total *= A->ext[i];
        gm.n++;
    }
    if (out->numel() != total) QB_FAIL(ctx, QB200_E_INVALID, "select_mode: output has wrong size");
    return gather(ctx, A, out, gm, index * st[mode_pos], 0);
}

int32_t qb200_conj(qb200_ctx* ctx, const qb200_tensor* A, qb200_tensor* out) {
    if (!A || !out || A->numel() != out->numel() || A->dtype != out->dtype)
        QB_FAIL(ctx, QB200_E_INVALID, "conj: bad argument");
    GatherModes gm;
    gm.n = 1;
    gm.ext[0] = A->numel();
    gm.stride[0] = 1;
    return gather(ctx, A, out, gm, 0, 1);
}

// out mode i = A mode perm[i]
int32_t qb200_permute(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* perm, qb200_tensor* out) {
    if (!A || !out || !perm || A->numel() != out->numel() || A->dtype != out->dtype)
        QB_FAIL(ctx, QB200_E_INVALID, "permute: bad argument");
    GatherModes gm;
    gm.n = A->rank;
    int64_t st[QB200_MAX_RANK];
    dense_strides(A, st);
    bool seen[QB200_MAX_RANK] = {false};
    for (int i = 0; i < A->rank; ++i) {
        int pidx = perm[i];
        if (pidx < 0 || pidx >= A->rank || seen[pidx]) QB_FAIL(ctx, QB200_E_INVALID, "permute: not a permutation");
        seen[pidx] = true;
        gm.ext[i] = A->ext[pidx];
        gm.stride[i] = st[pidx];
    }
    return gather(ctx, A, out, gm, 0, 0);
}

int32_t qb200_norm2(qb200_ctx* ctx, const qb200_tensor* A, double* result) {
    if (!A || !result) QB_FAIL(ctx, QB200_E_INVALID, "norm2: bad argument");
    double ss = 0.0;
    if (A->dtype == QB200_C64) {
        QB_TRY(qb_sumsq_f32(ctx, (const float*)A->data, A->numel() * 2, &ss));
    } else {
        int64_t nd = A->numel() * (A->dtype == QB200_C128 ? 2 : 1);
        QB_TRY(qb_sumsq(ctx, (const double*)A->data, nd, &ss));
    }
    *result = sqrt(ss);
    return QB200_OK;
}

__global__ void scale_all_kernel(c128* x, int64_t n, c128 f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] = cmul(x[i], f);
}
__global__ void scale_all_c64_kernel(float2* x, int64_t n, c128 f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        c128 v = cmul(make_double2(x[i].x, x[i].y), f);
        x[i] = make_float2((float)v.x, (float)v.y);
    }
}
__global__ void scale_all_real_kernel(double* x, int64_t n, double f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        x[i] *= f;
}

int32_t qb200_scale(qb200_ctx* ctx, qb200_tensor* A, const double factor[2]) {
    if (!A || !factor) QB_FAIL(ctx, QB200_E_INVALID, "scale: bad argument");
    int64_t n = A->numel();
    if (n == 0) return QB200_OK;
    if (A->dtype == QB200_C128)
        scale_all_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>((c128*)A->data, n,
                                                                         make_double2(factor[0], factor[1]));
    else if (A->dtype == QB200_C64)
        scale_all_c64_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>((float2*)A->data, n,
                                                                             make_double2(factor[0], factor[1]));
    else
        scale_all_real_kernel<<<grid_for(ctx, n, 256), 256, 0, ctx->stream>>>((double*)A->data, n, factor[0]);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

}  // extern "C"
