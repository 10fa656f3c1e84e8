// tcgen05 / TMEM building block for the ComplexF32 path (DESIGN.md §7, round-2 plan): a self-checking TF32
// `tcgen05.mma.cta_group::1.kind::tf32` probe, hand-written PTX, no CUTLASS.
//
//   * operands in shared memory in the canonical K-major, no-swizzle ("interleave") UMMA layout, written with ordinary
//     st.shared (the layout the ComplexF32 loader will produce: it has to pass the data through registers anyway for
//     the hi/lo TF32 split and the [re | im] embedding, so TMA cannot be used for it) + fence.proxy.async;
//   * accumulator in tensor memory (tcgen05.alloc, 128 lanes x N columns of FP32), one elected thread issues the
//     MMAs, completion through tcgen05.commit -> mbarrier, read back with tcgen05.ld.32x32b;
//   * pass 0 checks D = A B^T against integers computed on the CUDA cores (operands are small integers, exact in
//     TF32), pass 1 measures the issue-bound MMA rate with the operands resident in shared memory.
//
// Byte offset of element (row r, k) of an operand tile with R rows (A: R = 128, B: R = N):
//     (k / 4) * LBO + (r / 8) * SBO + (r % 8) * 16 + (k % 4) * 4,   SBO = 128 B, LBO = R * 16 B
// i.e. 8-row x 16-byte core matrices, core matrices of one 4-wide K chunk contiguous along the rows.  One MMA consumes
// K = 8 TF32 = two K chunks; the descriptor of the next MMA starts 2 * LBO further.
#include "../common.cuh"

namespace {

constexpr int TC_M = 128, TC_KBLK = 32;  // K elements resident per tile: 4 MMAs of K = 8

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start_address [0,14) (>>4), leading_byte_offset [16,30) (>>4), stride_byte_offset
    // [32,46) (>>4), version [46,48) = 1 (Blackwell), base_offset [49,52) = 0, lbo_mode [52] = 0, layout_type [61,64) = 0
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ uint32_t make_idesc_tf32(int M, int N) {
    // cute::UMMA::InstrDescriptor: c_format [4,6) = 1 (F32), a_format [7,10) = 2 (TF32), b_format [10,13) = 2,
    // a_major [15] = b_major [16] = 0 (K-major), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// waits for the phase with the given parity; gives up after 2^22 polls (the caller reports a timeout instead of hanging)
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

__device__ __forceinline__ int a_val(int r, int k) { return (r + 2 * k) % 7 - 3; }
__device__ __forceinline__ int b_val(int n, int k) { return (3 * n + k) % 5 - 2; }

// status: 0 ok, 1 timeout waiting for the MMA commit; maxerr: largest |D - expected| (bit pattern, atomicMax)
template <int N>
__global__ void __launch_bounds__(128, 1) tc5_probe_kernel(int iters, int check, unsigned* __restrict__ maxerr,
                                                           int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* As = reinterpret_cast<float*>(smem_raw);                    // 128 x 32 TF32 = 16 KB
    float* Bs = As + TC_M * TC_KBLK;                                   // N x 32
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t SBO = 128, LBO_A = TC_M * 16, LBO_B = N * 16;

    for (int e = tid; e < TC_M * TC_KBLK; e += 128) {
        int r = e % TC_M, k = e / TC_M;
        As[((k >> 2) * LBO_A + (r >> 3) * SBO + (r & 7) * 16 + (k & 3) * 4) >> 2] = (float)a_val(r, k);
    }
    for (int e = tid; e < N * TC_KBLK; e += 128) {
        int n = e % N, k = e / N;
        Bs[((k >> 2) * LBO_B + (n >> 3) * SBO + (n & 7) * 16 + (k & 3) * 4) >> 2] = (float)b_val(n, k);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // one warp allocates the accumulator columns (power of two >= 32) and later frees them
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "n"(N)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // generic-proxy writes of the operands -> visible to the async proxy the tensor core reads through
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_tf32(TC_M, N);
        const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < TC_KBLK / 8; ++kk) {
                const uint64_t da = make_smem_desc(a0 + kk * 2 * LBO_A, LBO_A, SBO);
                const uint64_t db = make_smem_desc(b0 + kk * 2 * LBO_B, LBO_B, SBO);
                umma_tf32(tmem_d, da, db, idesc, (it | kk) ? 1u : 0u);
            }
        }
        // arrives on the mbarrier when every MMA issued above has completed (implies fence::before_thread_sync)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                     : "memory");
    }
    const bool done = mbar_wait(smem_u32(&mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!done && tid == 0) atomicExch(status, 1);

    if (check && done) {
        // warp w owns TMEM lanes 32 w .. 32 w + 31 (= rows of D); each thread reads 32 consecutive columns of its row
        const int row = warp * 32 + lane;
        float worst = 0.f;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int expect = 0;
                for (int k = 0; k < TC_KBLK; ++k) expect += a_val(row, k) * b_val(c0 + j, k);
                worst = fmaxf(worst, fabsf(__uint_as_float(v[j]) - (float)(expect * iters)));
            }
        }
        atomicMax(maxerr, __float_as_uint(worst));  // non-negative floats order like their bit patterns
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(N) : "memory");
}

// ---- the same probe on the INT8 tensor pipe (kind::i8, S8 x S8 -> S32 in TMEM): the building block of the FP64
// ---- emulation by Ozaki splitting planned for the Jacobi update / Gram kernels.  One MMA consumes
// ---- K = 32 int8 = two 16-byte K chunks; the canonical layout is the same with 16 elements per 16-byte unit.
constexpr int TI_KBLK = 128;  // int8 K elements resident per tile: 4 MMAs of K = 32

__device__ __forceinline__ uint32_t make_idesc_i8(int M, int N) {
    // c_format [4,6) = 2 (S32), a_format [7,10) = b_format [10,13) = 1 (signed 8 bit), K-major, n_dim, m_dim
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

template <int N>
__global__ void __launch_bounds__(128, 1) tc5_probe_i8_kernel(int iters, int check, unsigned* __restrict__ maxerr,
                                                              int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    signed char* As = reinterpret_cast<signed char*>(smem_raw);  // 128 x 128 int8 = 16 KB
    signed char* Bs = As + TC_M * TI_KBLK;                        // N x 128
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr uint32_t SBO = 128, LBO_A = TC_M * 16, LBO_B = N * 16;

    for (int e = tid; e < TC_M * TI_KBLK; e += 128) {
        int r = e % TC_M, k = e / TC_M;
        As[(k >> 4) * LBO_A + (r >> 3) * SBO + (r & 7) * 16 + (k & 15)] = (signed char)a_val(r, k);
    }
    for (int e = tid; e < N * TI_KBLK; e += 128) {
        int n = e % N, k = e / N;
        Bs[(k >> 4) * LBO_B + (n >> 3) * SBO + (n & 7) * 16 + (k & 15)] = (signed char)b_val(n, k);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "n"(N)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;

    if (tid == 0) {
        const uint32_t idesc = make_idesc_i8(TC_M, N);
        const uint32_t a0 = smem_u32(As), b0 = smem_u32(Bs);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int kk = 0; kk < TI_KBLK / 32; ++kk) {
                const uint64_t da = make_smem_desc(a0 + kk * 2 * LBO_A, LBO_A, SBO);
                const uint64_t db = make_smem_desc(b0 + kk * 2 * LBO_B, LBO_B, SBO);
                umma_i8(tmem_d, da, db, idesc, (it | kk) ? 1u : 0u);
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar))
                     : "memory");
    }
    const bool done = mbar_wait(smem_u32(&mbar), 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!done && tid == 0) atomicExch(status, 1);

    if (check && done) {
        const int row = warp * 32 + lane;
        unsigned worst = 0;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                int expect = 0;
                for (int k = 0; k < TI_KBLK; ++k) expect += a_val(row, k) * b_val(c0 + j, k);
                int diff = (int)v[j] - expect * iters;
                worst = max(worst, (unsigned)(diff < 0 ? -diff : diff));
            }
        }
        atomicMax(maxerr, worst);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(N) : "memory");
}

template <int N>
int32_t run_probe_i8(qb200_ctx* ctx, int blocks, int iters, int check, unsigned* maxerr, int* status, double* ms) {
    const size_t smem = (size_t)(TC_M + N) * TI_KBLK + 1024;
    QB_CUDA(ctx, cudaFuncSetAttribute(tc5_probe_i8_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QB_TRY(qb200_timer_begin(ctx));
    tc5_probe_i8_kernel<N><<<blocks, 128, smem, ctx->stream>>>(iters, check, maxerr, status);
    QB_LAUNCH_CHECK(ctx);
    QB_TRY(qb200_timer_end(ctx, ms));
    return QB200_OK;
}

template <int N>
int32_t run_probe(qb200_ctx* ctx, int blocks, int iters, int check, unsigned* maxerr, int* status, double* ms) {
    const size_t smem = (size_t)(TC_M + N) * TC_KBLK * sizeof(float) + 1024;
    QB_CUDA(ctx, cudaFuncSetAttribute(tc5_probe_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    QB_TRY(qb200_timer_begin(ctx));
    tc5_probe_kernel<N><<<blocks, 128, smem, ctx->stream>>>(iters, check, maxerr, status);
    QB_LAUNCH_CHECK(ctx);
    QB_TRY(qb200_timer_end(ctx, ms));
    return QB200_OK;
}

}  // namespace

// out[0] = max |D - expected| of the checked pass (N = 128 and N = 256, 3 accumulation rounds; must be 0: the operands
// are small integers), out[1] / out[2] = issue-bound TF32 TFLOP/s with M = 128 and N = 128 / 256, operands resident in
// shared memory.  Returns QB200_E_CUDA with "timeout" in the message when the MMA completion never arrives.
extern "C" int32_t qb200_bench_tcgen05_tf32(qb200_ctx* ctx, double* out3) {
    if (!ctx || !out3) return QB200_E_INVALID;
    Workspace ws(ctx);
    unsigned* maxerr = ws.get<unsigned>(2);
    if (!maxerr) QB_FAIL(ctx, QB200_E_CUDA, "tcgen05 probe: workspace allocation failed");
    int* status = reinterpret_cast<int*>(maxerr + 1);
    QB_CUDA(ctx, cudaMemsetAsync(maxerr, 0, 2 * sizeof(unsigned), ctx->stream));
    double ms = 0.0;
    QB_TRY(run_probe<128>(ctx, 4, 3, 1, maxerr, status, &ms));
    QB_TRY(run_probe<256>(ctx, 4, 3, 1, maxerr, status, &ms));
    unsigned host[2] = {0, 0};
    QB_CUDA(ctx, cudaMemcpyAsync(host, maxerr, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (host[1] != 0) QB_FAIL(ctx, QB200_E_CUDA, "tcgen05 probe: timeout waiting for the MMA commit");
    float err;
    memcpy(&err, &host[0], sizeof(float));
    out3[0] = err;
    const int blocks = ctx->sm_count, iters = 20000;
    for (int which = 0; which < 2; ++which) {
        double best = 0.0;
        for (int rep = 0; rep < 3; ++rep) {
            if (which == 0)
                QB_TRY(run_probe<128>(ctx, blocks, iters, 0, maxerr, status, &ms));
            else
                QB_TRY(run_probe<256>(ctx, blocks, iters, 0, maxerr, status, &ms));
            const int N = which ? 256 : 128;
            double flops = (double)blocks * iters * (TC_KBLK / 8) * 2.0 * TC_M * N * 8;
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        out3[1 + which] = best;
    }
    QB_CUDA(ctx, cudaMemcpyAsync(host, maxerr, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (host[1] != 0) QB_FAIL(ctx, QB200_E_CUDA, "tcgen05 probe: timeout waiting for the MMA commit (timing pass)");
    return QB200_OK;
}

// INT8 twin (S8 x S8 -> S32, exact): out3 = {max |D - expected| (must be 0), issue-bound TOP/s at N = 128, at N = 256}.
// NOT yet run on hardware (written after the round's GPU budget was spent): first thing to run next round
// (tools/tc5_probe.py --i8).
extern "C" int32_t qb200_bench_tcgen05_i8(qb200_ctx* ctx, double* out3) {
    if (!ctx || !out3) return QB200_E_INVALID;
    Workspace ws(ctx);
    unsigned* maxerr = ws.get<unsigned>(2);
    if (!maxerr) QB_FAIL(ctx, QB200_E_CUDA, "tcgen05 i8 probe: workspace allocation failed");
    int* status = reinterpret_cast<int*>(maxerr + 1);
    QB_CUDA(ctx, cudaMemsetAsync(maxerr, 0, 2 * sizeof(unsigned), ctx->stream));
    double ms = 0.0;
    QB_TRY(run_probe_i8<128>(ctx, 4, 3, 1, maxerr, status, &ms));
    QB_TRY(run_probe_i8<256>(ctx, 4, 3, 1, maxerr, status, &ms));
    unsigned host[2] = {0, 0};
    QB_CUDA(ctx, cudaMemcpyAsync(host, maxerr, sizeof(host), cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (host[1] != 0) QB_FAIL(ctx, QB200_E_CUDA, "tcgen05 i8 probe: timeout waiting for the MMA commit");
    out3[0] = (double)host[0];
    const int blocks = ctx->sm_count, iters = 20000;
    for (int which = 0; which < 2; ++which) {
        double best = 0.0;
        for (int rep = 0; rep < 3; ++rep) {
            if (which == 0)
                QB_TRY(run_probe_i8<128>(ctx, blocks, iters, 0, maxerr, status, &ms));
            else
                QB_TRY(run_probe_i8<256>(ctx, blocks, iters, 0, maxerr, status, &ms));
            const int N = which ? 256 : 128;
            double ops = (double)blocks * iters * (TI_KBLK / 32) * 2.0 * TC_M * N * 32;
            best = std::max(best, ops / (ms * 1e-3) / 1e12);
        }
        out3[1 + which] = best;
    }
    return QB200_OK;
}
