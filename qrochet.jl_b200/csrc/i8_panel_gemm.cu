// FP64 complex panel product on the INT8 tensor pipe (Ozaki splitting):  C (M x 64) = A (M x 64) . B (64 x 64), all
// ComplexF64 column-major -- the shape of the Jacobi update X_p <- X_p W_p.
//
// STATUS: written after this round's GPU budget was spent -- compiles for sm_100a, NOT yet run on hardware.  It is
// reachable only through qb200_i8_panel_gemm (tools/i8_panel_check.py) and, with QB200_UPDATE_I8=1, as the update step of
// the Jacobi SVD (qb_i8_jacobi_update; off by default).  The arithmetic it has to
// reproduce BIT FOR BIT is tools/exp_ozaki.py::ozaki_complex (every floating-point operation below is exact or is
// performed in the same order there), which is how it is to be validated next round.
//
// Scheme (8 slices of 7 bits; K = 64):
//   * three real products (3M): P = Ar Br, Q = Ai Bi, S = (Ar + Ai)(Br + Bi);  Cr = P - Q,  Ci = S - P - Q;
//   * per product, every ROW of the A-side matrix and every COLUMN of the B-side matrix gets one exponent e (frexp
//     exponent of the largest magnitude + 1) and is cut into slices s_t = rint(r 2^7), r <- r 2^7 - s_t, starting from
//     r = a 2^-e: |s_t| <= 64 fits a signed byte, every step is exact in FP64;
//   * slice planes in shared memory in the canonical K-major no-swizzle UMMA layout (16 int8 per 16-byte unit);
//   * for each order d = 0..7 the pairs (i, d - i) are accumulated EXACTLY in INT32 in tensor memory by
//     tcgen05.mma.kind::i8 (M = 128, N = 64, two K = 32 instructions per pair), orders alternate between two TMEM
//     accumulators so that the epilogue of order d (tcgen05.ld, x 2^(-7(d+2) + e_row + e_col), add into the FP64
//     register accumulators) overlaps the MMAs of order d + 1.
// One CTA = 128 rows.  256 threads: slicing and epilogue by everyone, MMA issue by thread 0.
#include "common.cuh"

namespace {

constexpr int I8_M = 128, I8_N = 64, I8_K = 64, I8_NS = 8, I8_THREADS = 256;
constexpr uint32_t I8_SBO = 128, I8_LBO_A = I8_M * 16, I8_LBO_B = I8_N * 16;
constexpr uint32_t I8_APLANE = (I8_K / 16) * I8_LBO_A;  // 8 KB
constexpr uint32_t I8_BPLANE = (I8_K / 16) * I8_LBO_B;  // 4 KB
constexpr size_t I8_SMEM = (size_t)I8_NS * I8_APLANE + (size_t)3 * I8_NS * I8_BPLANE + 1024;  // 64 KB + 96 KB + slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((I8_SBO >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// frexp exponent + 1 of a positive finite double (0 for 0): x 2^-e has magnitude < 1/2
__device__ __forceinline__ int scale_exponent(double amax) {
    if (!(amax > 0.0)) return 0;
    int e;
    frexp(amax, &e);
    return e + 1;
}

// 2^e as a double, |e| small enough for a normal number
__device__ __forceinline__ double pow2i(int e) { return __hiloint2double((e + 1023) << 20, 0); }

// Cuts NV values (already scaled to |r| < 1/2) into 8 slices and writes, for every slice, NV consecutive int8 along K
// (NV = 32: two 16-byte units of the A layout; NV = 16: one unit of the B layout).
template <int NV>
__device__ __forceinline__ void slice_and_store(double (&r)[NV], unsigned char* plane0, uint32_t plane_bytes,
                                                uint32_t unit_off, uint32_t lbo) {
#pragma unroll 1
    for (int t = 0; t < I8_NS; ++t) {
        uint32_t packed[NV / 4];
#pragma unroll
        for (int c4 = 0; c4 < NV / 4; ++c4) {
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double x = r[c4 * 4 + b] * 128.0;
                const double s = rint(x);  // ties to even, |s| <= 64
                r[c4 * 4 + b] = x - s;
                w |= ((uint32_t)(int)s & 0xffu) << (8 * b);
            }
            packed[c4] = w;
        }
        unsigned char* dst = plane0 + (size_t)t * plane_bytes + unit_off;
#pragma unroll
        for (int u = 0; u < NV / 16; ++u)
            *reinterpret_cast<uint4*>(dst + u * lbo) =
                make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
    }
}

// circle-method round robin of svd_jacobi.cu (block pair of `pair` at `step`; nb even)
__device__ __forceinline__ void i8_rr_pair(int n, int step, int k, int& p, int& q) {
    if (n == 2) {
        p = 0;
        q = 1;
        return;
    }
    int a, b;
    if (k == 0) {
        a = n - 1;
        b = step;
    } else {
        a = (step + k) % (n - 1);
        b = (step - k + (n - 1)) % (n - 1);
    }
    p = min(a, b);
    q = max(a, b);
}

// JACOBI = false: C (M x 64) = A (M x 64) B, plain column-major panels (blockIdx.y unused).
// JACOBI = true : the update step of the one-sided block Jacobi, in place: A = C = X (ld lda), blockIdx.y = pair, the
//                 64 panel columns are the two 32-column blocks (I, J) of the pair, B = Wg + pair * 64 * 64, pairs with
//                 flags[pair] == 0 (W = identity) are skipped.  In place is safe: a CTA reads its 128 rows of the panel
//                 for all three products before it writes them, and no other CTA touches them.
template <bool JACOBI>
__global__ void __launch_bounds__(I8_THREADS, 1)
    i8_panel_gemm_kernel(const c128* A, int64_t lda, int M, const c128* __restrict__ B, c128* C, int64_t ldc,
                         int* __restrict__ status, int nb, int step, const int* __restrict__ flags) {  // A may alias C
    int blkI = 0, blkJ = 1;
    if (JACOBI) {
        if (!flags[blockIdx.y]) return;  // uniform for the CTA, before anything is allocated
        i8_rr_pair(nb, step, (int)blockIdx.y, blkI, blkJ);
        B += (size_t)blockIdx.y * (I8_K * I8_N);
    }
    auto col_of = [&](int c) -> int64_t {  // panel column c -> column of the matrix
        if (!JACOBI) return c;
        return (c < 32) ? (int64_t)blkI * 32 + c : (int64_t)blkJ * 32 + (c - 32);
    };
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = smem;                                  // 8 planes of the current product
    unsigned char* sB = smem + (size_t)I8_NS * I8_APLANE;      // 3 products x 8 planes
    __shared__ __align__(8) uint64_t mbar_full[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ double part_max[2][I8_M];      // partial row maxima (two column halves) / column maxima (four k quarters)
    __shared__ double part_maxb[4][I8_N];
    __shared__ int ea_s[I8_M];                // row exponents of the current product
    __shared__ int eb_s[3][I8_N];             // column exponents per product
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * I8_M;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_full[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_full[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // two INT32 accumulators of 64 columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_smem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }

    // ---- B side, once: thread (n = tid % 64, kq = tid / 64) holds W[kq*16 .. +15][n] ----
    const int bn = tid & 63, bkq = tid >> 6;
    c128 bw[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) bw[j] = B[(bkq * 16 + j) + (int64_t)bn * I8_K];
    for (int prod = 0; prod < 3; ++prod) {
        double r[16], mx = 0.0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            r[j] = prod == 0 ? bw[j].x : (prod == 1 ? bw[j].y : bw[j].x + bw[j].y);
            mx = fmax(mx, fabs(r[j]));
        }
        part_maxb[bkq][bn] = mx;
        __syncthreads();
        const int e = scale_exponent(fmax(fmax(part_maxb[0][bn], part_maxb[1][bn]), fmax(part_maxb[2][bn], part_maxb[3][bn])));
        if (bkq == 0) eb_s[prod][bn] = e;
        const double sc = pow2i(-e);
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] *= sc;
        // B operand row n, K unit bkq: offset inside a plane
        slice_and_store<16>(r, sB + (size_t)prod * I8_NS * I8_BPLANE, I8_BPLANE,
                            (uint32_t)(bkq * I8_LBO_B + (bn >> 3) * I8_SBO + (bn & 7) * 16), I8_LBO_B);
        __syncthreads();  // part_maxb is reused by the next product
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;

    // ---- A side: thread (row = tid % 128, half = tid / 128) holds 32 consecutive columns of its row ----
    const int arow = tid & 127, ahalf = tid >> 7;
    const bool row_ok = (m0 + arow) < M;
    // epilogue mapping: TMEM lane quarter q (rows), 32-column half hsel
    const int q = warp & 3, hsel = warp >> 2;
    const int erow = q * 32 + lane;
    double acc_re[32], acc_im[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc_re[j] = acc_im[j] = 0.0;
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(I8_N >> 3) << 17) | ((uint32_t)(I8_M >> 4) << 24);
    int uses[2] = {0, 0};  // completed-phase counters of the two TMEM-full barriers (uniform across the CTA)
    bool alive = true;

    for (int prod = 0; prod < 3; ++prod) {
        {
            double r[32], mx = 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                c128 v = make_double2(0.0, 0.0);
                if (row_ok) v = A[(m0 + arow) + col_of(ahalf * 32 + c) * lda];
                r[c] = prod == 0 ? v.x : (prod == 1 ? v.y : v.x + v.y);
                mx = fmax(mx, fabs(r[c]));
            }
            part_max[ahalf][arow] = mx;
            __syncthreads();  // also: every MMA of the previous product has completed (all its commits were waited on)
            const int e = scale_exponent(fmax(part_max[0][arow], part_max[1][arow]));
            if (ahalf == 0) ea_s[arow] = e;
            const double sc = pow2i(-e);
#pragma unroll
            for (int c = 0; c < 32; ++c) r[c] *= sc;
            // A operand row arow, K units 2 * ahalf and 2 * ahalf + 1
            slice_and_store<32>(r, sA, I8_APLANE,
                                (uint32_t)(ahalf * 2 * I8_LBO_A + (arow >> 3) * I8_SBO + (arow & 7) * 16), I8_LBO_A);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();

        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB + (size_t)prod * I8_NS * I8_BPLANE);
        auto issue_order = [&](int d) {  // thread 0: all slice pairs (i, d - i) into accumulator d & 1, then commit
            const uint32_t dcol = tmem_d + (uint32_t)((d & 1) * I8_N);
            for (int i = 0; i <= d; ++i) {
                const int j = d - i;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                    const uint64_t da = umma_desc(a0 + i * I8_APLANE + kk * 2 * I8_LBO_A, I8_LBO_A);
                    const uint64_t db = umma_desc(b0 + j * I8_BPLANE + kk * 2 * I8_LBO_B, I8_LBO_B);
                    umma_i8(dcol, da, db, idesc, (i | kk) ? 1u : 0u);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&mbar_full[d & 1]))
                         : "memory");
        };
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_order(0);
            issue_order(1);
        }
        const double row_scale = pow2i(ea_s[erow]);
        for (int d = 0; d < I8_NS; ++d) {
            const int b = d & 1;
            alive = mbar_wait(smem_u32(&mbar_full[b]), (uint32_t)(uses[b] & 1)) && alive;
            uses[b]++;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t v[32];
            tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * I8_N + hsel * 32), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const double ds = pow2i(-7 * (d + 2)) * row_scale;  // exact power of two
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const double t = (double)(int)v[j] * ds * pow2i(eb_s[prod][hsel * 32 + j]);  // exact: integer x 2^e
                if (prod == 0) {
                    acc_re[j] += t;
                    acc_im[j] -= t;
                } else if (prod == 1) {
                    acc_re[j] -= t;
                    acc_im[j] -= t;
                } else {
                    acc_im[j] += t;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();  // every warp has read accumulator b: it may be overwritten by order d + 2
            if (tid == 0 && d + 2 < I8_NS) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                issue_order(d + 2);
            }
        }
    }
    if (!alive && tid == 0) *reinterpret_cast<volatile int*>(status) = 1;
    if (alive && (m0 + erow) < M) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            C[(m0 + erow) + col_of(hsel * 32 + j) * ldc] = make_double2(acc_re[j], acc_im[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");
}

}  // namespace

// C (M x 64, ldc) = A (M x 64, lda) . B (64 x 64, ld 64) on device pointers held by tensors; C must not alias A.
// Experimental entry point (see the header of this file).  Synchronises and reports an MMA-commit timeout.
extern "C" int32_t qb200_i8_panel_gemm(qb200_ctx* ctx, const qb200_tensor* A, const qb200_tensor* B, qb200_tensor* C) {
    if (!ctx || !A || !B || !C) QB_FAIL(ctx, QB200_E_INVALID, "i8_panel_gemm: null argument");
    if (A->dtype != QB200_C128 || B->dtype != QB200_C128 || C->dtype != QB200_C128 || A->rank != 2 || B->rank != 2 ||
        C->rank != 2 || A->ext[1] != I8_K || B->ext[0] != I8_K || B->ext[1] != I8_N || C->ext[0] != A->ext[0] ||
        C->ext[1] != I8_N || A->data == C->data)
        QB_FAIL(ctx, QB200_E_INVALID, "i8_panel_gemm: needs ComplexF64 A (M x 64), B (64 x 64), C (M x 64), C != A");
    const int64_t M = A->ext[0];
    if (M <= 0) return QB200_OK;
    if (M > (1 << 30)) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "i8_panel_gemm: too many rows");
    static bool attr_set = false;
    if (!attr_set) {
        QB_CUDA(ctx, cudaFuncSetAttribute(i8_panel_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8_SMEM));
        attr_set = true;
    }
    i8_panel_gemm_kernel<false><<<(unsigned)((M + I8_M - 1) / I8_M), I8_THREADS, I8_SMEM, ctx->stream>>>(
        (const c128*)A->data, M, (int)M, (const c128*)B->data, (c128*)C->data, M, qb_async_status(ctx), 2, 0, nullptr);
    QB_LAUNCH_CHECK(ctx);
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return qb_check_async_status(ctx);
}

// One update step of the block Jacobi on the INT8 pipe (experimental, QB200_UPDATE_I8=1 in svd_jacobi.cu): X_p <- X_p W_p
// for every active pair, in place.  Same arguments as jacobi_update_kernel.  Asynchronous; a timeout surfaces at the
// next synchronising call.  NOTE: no column scaling yet: fine for panels whose column norms are
// within ~1e6 of each other, not for strongly graded ones.
int32_t qb_i8_jacobi_update(qb200_ctx* ctx, c128* Z, int64_t ldz, int rows, int nb, int step, const c128* Wg,
                            const int* flags, int npairs) {
    static bool attr_set = false;
    if (!attr_set) {
        QB_CUDA(ctx, cudaFuncSetAttribute(i8_panel_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8_SMEM));
        attr_set = true;
    }
    dim3 grid((unsigned)((rows + I8_M - 1) / I8_M), (unsigned)npairs);
    i8_panel_gemm_kernel<true><<<grid, I8_THREADS, I8_SMEM, ctx->stream>>>(Z, ldz, rows, Wg, Z, ldz, qb_async_status(ctx), nb,
                                                                         step, flags);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}
