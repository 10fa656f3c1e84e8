// K2/K3/K6/K7/K8: HBM-bound helpers -- mode scale (with pinv), slice, select, conj, permute, norms,
// gate application on the two-site wave function.  All are single-pass, 16-byte vectorised
// (one ComplexF64 per access), grid-stride with grids sized as multiples of the SM count.
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

static int32_t gather(qb200_ctx* ctx, const qb200_tensor* A, qb200_tensor* out, const GatherModes& gm, int64_t base,
                      int conj) {
    int64_t total = 1;
    for (int j = 0; j < gm.n; ++j) total *= gm.ext[j];
    if (total == 0) return QB200_OK;
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
