// K1 for ComplexF32 on the 5th-generation tensor cores: permutation-fused complex GEMM with the 3xTF32 split on
// `tcgen05.mma.cta_group::1.kind::tf32`, accumulators in tensor memory.  Hand-written PTX (primitives validated by
// tc5_probe.cu).  Same GemmArgs contract as gemm_c64.cu / gemm_c128.cu.  Measured on B200: 109.8 TFLOP/s at 4096^3
// (mma.sync kernel: 61.4), 74.4 on the 12-mode permuted contraction (38.6), relative error 1.9e-6 at K = 4096.
//
// Real embedding of the complex product, per K block of 16 complex numbers (x = xh + xl, TF32 split):
//
//   [D_re | D_im] (128 x 128 FP32, TMEM)  +=  sum over terms (lh, hl, hh)   [A_r  A_i] (128 x 32)  .  | B_r   B_i |  (32 x 128)
//                                                                                                    | -B_i  B_r |
//
// Shared memory holds, per stage, A as four 16-wide K blocks {Ah_r, Ah_i, Al_r, Al_i} (128 rows) and B as four blocks
// {Bh, Bl} x {pairs with A_r, pairs with A_i} (128 rows = 64 real-part + 64 imaginary-part output columns), all in the
// canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices, LBO = 2048 B between 4-wide K chunks, SBO =
// 128 B between 8-row groups).  The SAME Ah chunks feed the hh and hl terms through two descriptors, so nothing is
// stored twice.  12 MMAs (M = 128, N = 128, K = 8) per stage.
//
// The operands have to pass through registers (hi/lo split, [re | im] embedding, the gather of the fused permutation),
// so the loader is ordinary code: every thread loads 4 consecutive k of one row, splits, and writes 16-byte vectors
// (conflict-free: consecutive lanes = consecutive rows = consecutive 16-byte slots of a core-matrix column), then
// fence.proxy.async + barrier, and ONE thread issues the MMAs of the stage and commits them to the stage's mbarrier,
// which is what the loaders wait on before they overwrite that stage (2 stages: the fill of one overlaps the MMAs
// of the other).  Epilogue: 8 warps read their TMEM lane quarter with tcgen05.ld.32x32b (re and im columns of 32
// output columns each), apply alpha / beta and write C through the offset tables (rows of C across lanes).
#include "gemm_c128.cuh"
#include "tc5.cuh"

#include <algorithm>
#include <cstdlib>

namespace qb {

namespace {

using namespace tc5;

constexpr int TM = 128, TN = 64, TBK = 16, T_THREADS = 256;
constexpr uint32_t T_STAGE_BYTES = 2 * T_OPER_BYTES;          // A + B
constexpr size_t GEMM_TC5_SMEM = 2 * T_STAGE_BYTES + 1024;    // 2 stages + alignment slack

__global__ void __launch_bounds__(T_THREADS, 1) gemm_c64_tc5_kernel(const GemmArgs p, int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
    __shared__ __align__(8) uint64_t mbar_free[2];
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int tile_m, tile_n;
    {  // grouped rasterisation, as in the other GEMM kernels
        const int gm = gridDim.x, gn = gridDim.y, GROUP = 8;
        const int id = blockIdx.x + gm * blockIdx.y;
        const int per_group = GROUP * gn;
        const int first_m = (id / per_group) * GROUP;
        const int gsz = min(gm - first_m, GROUP);
        tile_m = first_m + (id % per_group) % gsz;
        tile_n = (id % per_group) / gsz;
    }
    const int m0 = tile_m * TM, n0 = tile_n * TN;
    const int z = blockIdx.z;
    const float2* __restrict__ A = reinterpret_cast<const float2*>(p.A) + p.ab.at(z);
    const float2* __restrict__ B = reinterpret_cast<const float2*>(p.B) + p.bb.at(z);
    float2* __restrict__ C = reinterpret_cast<float2*>(p.C) + p.cb.at(z);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_free[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_free[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_smem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;

    // loader items: A (row, k group) x 2 per thread, B (column, k group) x 1 per thread
    const int a_row[2] = {tid & 127, tid & 127}, a_g[2] = {tid >> 7, 2 + (tid >> 7)};
    const int b_col = tid & 63, b_g = tid >> 6;
    const bool a_ok = (m0 + a_row[0]) < p.M, b_ok = (n0 + b_col) < p.N;
    const int64_t a_moff = a_ok ? p.am.at(m0 + a_row[0]) : 0;
    const int64_t b_noff = b_ok ? p.bn.at(n0 + b_col) : 0;
    const float sa = p.conjA ? -1.f : 1.f, sb = p.conjB ? -1.f : 1.f;
    const uint32_t a_slot = (uint32_t)((a_row[0] >> 3) * T_SBO + (a_row[0] & 7) * 16);   // inside a K chunk
    const uint32_t br_slot = (uint32_t)((b_col >> 3) * T_SBO + (b_col & 7) * 16);        // real-part output column
    const uint32_t bi_slot = (uint32_t)(((64 + b_col) >> 3) * T_SBO + ((64 + b_col) & 7) * 16);  // imaginary-part column

    const int KT = (p.K + TBK - 1) / TBK;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    bool alive = true;

    // Two-level accumulation.  tcgen05 accumulates into TMEM with round-toward-zero (measured: error ~ 0.3 ulp per
    // accumulating MMA, growing linearly with K), so the TMEM accumulator is closed every PROMOTE stages (K = 128):
    // every thread adds its 32 + 32 TMEM values to FP32 register accumulators (round to nearest) and the next MMA
    // starts the TMEM accumulator from zero.
    constexpr int PROMOTE = 8;
    const int q = warp & 3, hsel = warp >> 2;   // TMEM lane quarter (rows) / 32-column half of the 64 output columns
    const uint32_t t_re = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(hsel * 32);
    const uint32_t t_im = t_re + 64u;
    float acc_re[32], acc_im[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc_re[j] = acc_im[j] = 0.f;
    auto promote = [&]() {  // all MMAs issued so far have completed (the caller waited on their commits)
        uint32_t vr[32], vi[32];
        tmem_ld32(t_re, vr);
        tmem_ld32(t_im, vi);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            acc_re[j] += __uint_as_float(vr[j]);
            acc_im[j] += __uint_as_float(vi[j]);
        }
    };

    // global -> register prefetch of one stage (4 consecutive k of 2 A rows-items and 1 B item per thread)
    float2 pa[2][4], pb[4];
    auto prefetch = [&](int kt) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kt * TBK + a_g[i] * 4 + j;
                pa[i][j] = (a_ok && k < p.K) ? A[a_moff + p.ak.at(k)] : make_float2(0.f, 0.f);
            }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = kt * TBK + b_g * 4 + j;
            pb[j] = (b_ok && k < p.K) ? B[b_noff + p.bk.at(k)] : make_float2(0.f, 0.f);
        }
    };
    if (KT > 0) prefetch(0);

    for (int kt = 0; kt < KT; ++kt) {
        const int s = kt & 1;
        unsigned char* sA = smem + (size_t)s * T_STAGE_BYTES;
        unsigned char* sB = sA + T_OPER_BYTES;
        if (kt >= 2) alive = mbar_wait(smem_u32(&mbar_free[s]), (uint32_t)(((kt >> 1) + 1) & 1)) && alive;
        const bool closing = (kt > 0) && (kt % PROMOTE == 0);
        if (closing)  // the previous stage's MMAs too: then everything issued so far is in the accumulator
            alive = mbar_wait(smem_u32(&mbar_free[s ^ 1]), (uint32_t)(((kt - 1) >> 1) & 1)) && alive;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (closing && alive) promote();
        // ---- fill stage s from the prefetched registers ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float xr[4], xi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                xr[j] = pa[i][j].x;
                xi[j] = pa[i][j].y * sa;
            }
            float4 rh, rl, ih, il;
            split4(xr, rh, rl);
            split4(xi, ih, il);
            // K blocks of A: 0 = Ah_r, 1 = Ah_i, 2 = Al_r, 3 = Al_i; chunk = 4 * block + k group
            *reinterpret_cast<float4*>(sA + (0 * 4 + a_g[i]) * T_LBO + a_slot) = rh;
            *reinterpret_cast<float4*>(sA + (1 * 4 + a_g[i]) * T_LBO + a_slot) = ih;
            *reinterpret_cast<float4*>(sA + (2 * 4 + a_g[i]) * T_LBO + a_slot) = rl;
            *reinterpret_cast<float4*>(sA + (3 * 4 + a_g[i]) * T_LBO + a_slot) = il;
        }
        {
            float xr[4], xi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                xr[j] = pb[j].x;
                xi[j] = pb[j].y * sb;
            }
            float4 rh, rl, ih, il;
            split4(xr, rh, rl);
            split4(xi, ih, il);
            // K blocks of B: 0 = (Bh, pairs with A_r), 1 = (Bh, pairs with A_i), 2 = (Bl, A_r), 3 = (Bl, A_i)
            // real-part column n:      block(A_r) = B_r, block(A_i) = -B_i;  imaginary-part column 64 + n: B_i, B_r
            *reinterpret_cast<float4*>(sB + (0 * 4 + b_g) * T_LBO + br_slot) = rh;
            *reinterpret_cast<float4*>(sB + (1 * 4 + b_g) * T_LBO + br_slot) = neg4(ih);
            *reinterpret_cast<float4*>(sB + (0 * 4 + b_g) * T_LBO + bi_slot) = ih;
            *reinterpret_cast<float4*>(sB + (1 * 4 + b_g) * T_LBO + bi_slot) = rh;
            *reinterpret_cast<float4*>(sB + (2 * 4 + b_g) * T_LBO + br_slot) = rl;
            *reinterpret_cast<float4*>(sB + (3 * 4 + b_g) * T_LBO + br_slot) = neg4(il);
            *reinterpret_cast<float4*>(sB + (2 * 4 + b_g) * T_LBO + bi_slot) = il;
            *reinterpret_cast<float4*>(sB + (3 * 4 + b_g) * T_LBO + bi_slot) = rl;
        }
        if (kt + 1 < KT) prefetch(kt + 1);  // in flight across the barrier, the MMA issue and the next wait
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (tensor core)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();   // also: every warp has finished promote()'s TMEM reads before the accumulator is restarted
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            const int fresh = (kt % PROMOTE == 0) ? 1 : 0;   // first stage of a TMEM accumulation window
            // small terms first: lh (Al x Bh), hl (Ah x Bl), hh (Ah x Bh); blk = 0 (A_r part) / 1 (A_i part); two K = 8 halves
#pragma unroll
            for (int term = 0; term < 3; ++term) {
                const int ablk0 = (term == 0) ? 2 : 0;   // Al blocks 2,3 ; Ah blocks 0,1
                const int bblk0 = (term == 1) ? 2 : 0;   // Bl blocks 2,3 ; Bh blocks 0,1
#pragma unroll
                for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t da = umma_desc(a0 + ((ablk0 + blk) * 4 + 2 * half) * T_LBO);
                        const uint64_t db = umma_desc(b0 + ((bblk0 + blk) * 4 + 2 * half) * T_LBO);
                        umma_tf32(tmem_d, da, db, idesc, (fresh && term == 0 && blk == 0 && half == 0) ? 0u : 1u);
                    }
            }
            // arrives when the MMAs above (and all earlier ones) have completed: frees stage s for the loaders
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                             smem_u32(&mbar_free[s]))
                         : "memory");
        }
    }
    // every MMA completes in issue order: the last stage's commit covers the whole product
    if (KT > 0) alive = mbar_wait(smem_u32(&mbar_free[(KT - 1) & 1]), (uint32_t)(((KT - 1) >> 1) & 1)) && alive;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (!alive && tid == 0) *reinterpret_cast<volatile int*>(status) = 1;  // pinned host word, read at the next sync

    if (alive) {
        // warp w: TMEM lanes 32 (w % 4) .. +31 = rows; output columns 32 (w / 4) .. +31: re at TMEM column n, im at 64 + n
        const int m = m0 + q * 32 + lane;
        if (KT > 0) promote();   // the last (open) accumulation window
        if (m < p.M) {
            const int64_t mo = p.cm.at(m);
            const float alr = (float)p.alpha.x, ali = (float)p.alpha.y, ber = (float)p.beta.x, bei = (float)p.beta.y;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + hsel * 32 + j;
                if (n >= p.N) continue;
                float2* dst = C + mo + p.cn.at(n);
                const float re = acc_re[j], im = acc_im[j];
                float2 o = make_float2(alr * re - ali * im, alr * im + ali * re);
                if (!p.beta_zero) {
                    float2 old = *dst;
                    o.x += ber * old.x - bei * old.y;
                    o.y += ber * old.y + bei * old.x;
                }
                *dst = o;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_d) : "memory");
}

}  // namespace

int32_t init_gemm_c64_tc5(qb200_ctx* ctx) {
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c64_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_TC5_SMEM));
    return QB200_OK;
}

// QB200_C64_TCGEN05=0 keeps ComplexF32 contractions on the mma.sync kernel of gemm_c64.cu (A/B switch).  Default: the
// tcgen05 kernel -- it passes the ComplexF32 parity suites (tests/test_gpu_c64.py, tests/test_gpu_kernels.py: 106
// tests), is more accurate at long K (two-level accumulation) and 1.5-1.9x faster (profiles/r1b_tcgen05_c64_gemm.txt).
bool gemm_c64_tc5_enabled() {
    static const bool on = [] {
        const char* e = getenv("QB200_C64_TCGEN05");
        return !(e && e[0] == '0');
    }();
    return on;
}

// `args` already oriented (M >= N side on the 128-wide tile) by launch_gemm_c64; no split-K in this kernel
int32_t launch_gemm_c64_tc5(qb200_ctx* ctx, const GemmArgs& args) {
    dim3 grid((args.M + TM - 1) / TM, (args.N + TN - 1) / TN, args.batch);
    if (grid.y > 65535 || grid.z > 65535) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "gemm_c64_tc5 grid too large");
    int* status = qb_async_status(ctx);  // raised by the kernel on an mbarrier timeout, checked at the next sync
    gemm_c64_tc5_kernel<<<grid, T_THREADS, GEMM_TC5_SMEM, ctx->stream>>>(args, status);
    QB_LAUNCH_CHECK(ctx);
    static const bool check_now = getenv("QB200_C64_TCGEN05_CHECK") != nullptr;  // debugging aid: check after every launch
    if (check_now) {
        QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        QB_TRY(qb_check_async_status(ctx));
    }
    return QB200_OK;
}

}  // namespace qb
