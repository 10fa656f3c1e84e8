// Low-precision stage of the mixed-precision Jacobi SVD (svd_jacobi.cu, "stage A"): the update S_p <- S_p W_p of the
// FP32 shadow S = [X32; V32] on the 5th-generation tensor cores -- tcgen05.mma kind::tf32 with the 3xTF32 split (FP32
// accuracy), accumulators in tensor memory.  Same real embedding and operand layout as gemm_c64_tc5.cu:
//
//   [D_re | D_im] (128 rows x 128, TMEM)  =  sum over (lh, hl, hh)   [A_r  A_i] (128 x 128)  .  | W_r   W_i |  (128 x 128)
//                                                                                             | -W_i  W_r |
//
// A CTA owns one block pair and a set of 128-row chunks of S.  The whole W_p (64 x 64 complex = 4 K blocks of 16) is
// split, embedded and parked in shared memory ONCE (128 KB); the chunks stream through a 2-stage ring of A operands
// (32 KB each, one K block per stage): global -> registers (coalesced: lanes = consecutive rows of a column) -> hi / lo
// split -> canonical K-major core matrices.  One thread issues 12 MMAs (M = 128, N = 128, K = 8) per stage, 48 per
// chunk in ONE TMEM accumulation window (the window of gemm_c64_tc5.cu is 96), commits to the stage's mbarrier.  Two
// TMEM accumulators alternate: the epilogue of a chunk (all warps read their TMEM lane quarter and store the chunk back
// IN PLACE -- its rows of the 64 panel columns were consumed into shared memory long before, no other CTA touches them)
// runs while the tensor core works on the first K block of the next chunk.  The global loads run a whole chunk ahead in
// registers (64 KB in flight per SM): the kernel is a stream over S, not a GEMM.
#include <algorithm>

#include "common.cuh"
#include "jacobi_rr.cuh"
#include "tc5.cuh"

namespace qb {

namespace {

using namespace tc5;

constexpr int LU_ROWS = 128, LU_THREADS = 256, LU_KBLK = 16, LU_NKB = JP / LU_KBLK;
constexpr uint32_t LU_B_BYTES = LU_NKB * T_OPER_BYTES;  // W_p, all K blocks: 128 KB
constexpr size_t LU_SMEM = LU_B_BYTES + 2 * T_OPER_BYTES + 1024;

__global__ void __launch_bounds__(LU_THREADS, 1)
    lp_update_tc5_kernel(float2* __restrict__ S, int64_t lds, int nchunks, int nb, int step, const c128* __restrict__ Wg,
                         const int* __restrict__ flags, int* __restrict__ status) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t mbar_free[2];
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pair = blockIdx.y;
    if (!flags[pair]) return;  // the pair was orthogonal: W_p = I (uniform over the CTA, before any allocation)
    int I, J;
    rr_pair(nb, step, pair, I, J);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_free[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_free[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base_smem))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_smem;

    unsigned char* sBall = smem;
    unsigned char* sAring = smem + LU_B_BYTES;

    // ---- W_p -> B operand, all four K blocks (thread: output column n = tid & 63, k group of 4 = tid >> 6) ----
    {
        const int b_col = tid & 63, b_g = tid >> 6;
        const uint32_t br_slot = (uint32_t)((b_col >> 3) * T_SBO + (b_col & 7) * 16);                 // real-part output column
        const uint32_t bi_slot = (uint32_t)(((64 + b_col) >> 3) * T_SBO + ((64 + b_col) & 7) * 16);   // imaginary-part column
        const c128* wsrc = Wg + (size_t)pair * (JP * JP) + (size_t)b_col * JP;  // W(k, n) at n * 64 + k
#pragma unroll
        for (int kb = 0; kb < LU_NKB; ++kb) {
            float xr[4], xi[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const c128 w = wsrc[kb * LU_KBLK + b_g * 4 + j];
                xr[j] = (float)w.x;
                xi[j] = (float)w.y;
            }
            float4 rh, rl, ih, il;
            split4(xr, rh, rl);
            split4(xi, ih, il);
            unsigned char* sB = sBall + (size_t)kb * T_OPER_BYTES;
            // K blocks of B: 0 = (Bh, pairs with A_r), 1 = (Bh, pairs with A_i), 2 = (Bl, A_r), 3 = (Bl, A_i)
            *reinterpret_cast<float4*>(sB + (0 * 4 + b_g) * T_LBO + br_slot) = rh;
            *reinterpret_cast<float4*>(sB + (1 * 4 + b_g) * T_LBO + br_slot) = neg4(ih);
            *reinterpret_cast<float4*>(sB + (0 * 4 + b_g) * T_LBO + bi_slot) = ih;
            *reinterpret_cast<float4*>(sB + (1 * 4 + b_g) * T_LBO + bi_slot) = rh;
            *reinterpret_cast<float4*>(sB + (2 * 4 + b_g) * T_LBO + br_slot) = rl;
            *reinterpret_cast<float4*>(sB + (3 * 4 + b_g) * T_LBO + br_slot) = neg4(il);
            *reinterpret_cast<float4*>(sB + (2 * 4 + b_g) * T_LBO + bi_slot) = il;
            *reinterpret_cast<float4*>(sB + (3 * 4 + b_g) * T_LBO + bi_slot) = rl;
        }
    }

    // A loader items: (row = tid & 127, k group of 4) x 2 per thread
    const int a_row = tid & 127;
    const int a_g[2] = {tid >> 7, 2 + (tid >> 7)};
    const uint32_t a_slot = (uint32_t)((a_row >> 3) * T_SBO + (a_row & 7) * 16);
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int q = warp & 3, hsel = warp >> 2;  // TMEM lane quarter (rows) / 32-column half of the 64 output columns
    const uint32_t t_re = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(hsel * 32);
    const uint32_t t_im = t_re + 64u;

    // register prefetch, one whole chunk ahead: pa[kb] holds K block kb of the chunk being processed until it is written
    // to shared memory, then K block kb of the NEXT chunk (64 KB in flight per SM: the loads have a chunk time to land)
    float2 pa[LU_NKB][2][4];
    auto prefetch = [&](int chunk, int kb, float2 (&dst)[2][4]) {
        const float2* src = S + (int64_t)chunk * LU_ROWS + a_row;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[i][j] = src[panel_col(I, J, kb * LU_KBLK + a_g[i] * 4 + j) * lds];
    };
    bool alive = true;
    // chunk (rows) <- TMEM accumulator `buf`, once the commit of running stage `tl` (the chunk's last) has arrived
    auto epilogue = [&](int chunk, int buf, int tl) {
        alive = mbar_wait(smem_u32(&mbar_free[tl & 1]), (uint32_t)((tl >> 1) & 1)) && alive;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (alive) {
            uint32_t vr[32], vi[32];
            tmem_ld32(t_re + (uint32_t)(buf * 128), vr);
            tmem_ld32(t_im + (uint32_t)(buf * 128), vi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float2* dst = S + (int64_t)chunk * LU_ROWS + q * 32 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                dst[panel_col(I, J, hsel * 32 + j) * lds] = make_float2(__uint_as_float(vr[j]), __uint_as_float(vi[j]));
        }
    };

    int t = 0;  // running stage counter (never reset: it carries the mbarrier phases)
    int chunk = blockIdx.x;
    if (chunk < nchunks) {
#pragma unroll
        for (int kb = 0; kb < LU_NKB; ++kb) prefetch(chunk, kb, pa[kb]);
    }
    int prev_chunk = -1, prev_tl = 0, ci = 0;
    for (; chunk < nchunks; chunk += gridDim.x, ++ci) {
        const int next = chunk + (int)gridDim.x;
        const uint32_t acc = tmem_d + (uint32_t)((ci & 1) * 128);  // two accumulators: the epilogue of a chunk overlaps the next one
#pragma unroll
        for (int kb = 0; kb < LU_NKB; ++kb) {
            const int s = t & 1;
            unsigned char* sA = sAring + (size_t)s * T_OPER_BYTES;
            if (t >= 2) alive = mbar_wait(smem_u32(&mbar_free[s]), (uint32_t)(((t >> 1) + 1) & 1)) && alive;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float xr[4], xi[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    xr[j] = pa[kb][i][j].x;
                    xi[j] = pa[kb][i][j].y;
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
            if (next < nchunks) prefetch(next, kb, pa[kb]);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();  // also: every warp finished the TMEM reads of the chunk before last (same accumulator)
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sBall + (size_t)kb * T_OPER_BYTES);
                // small terms first: lh (Al x Bh), hl (Ah x Bl), hh (Ah x Bh); blk = A_r / A_i part; two K = 8 halves
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const int ablk0 = (term == 0) ? 2 : 0;
                    const int bblk0 = (term == 1) ? 2 : 0;
#pragma unroll
                    for (int blk = 0; blk < 2; ++blk)
#pragma unroll
                        for (int half = 0; half < 2; ++half) {
                            const uint64_t da = umma_desc(a0 + ((ablk0 + blk) * 4 + 2 * half) * T_LBO);
                            const uint64_t db = umma_desc(b0 + ((bblk0 + blk) * 4 + 2 * half) * T_LBO);
                            umma_tf32(acc, da, db, idesc, (kb == 0 && term == 0 && blk == 0 && half == 0) ? 0u : 1u);
                        }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                                 smem_u32(&mbar_free[s]))
                             : "memory");
            }
            ++t;
            // the previous chunk's result leaves while the tensor core works on this chunk's first K block (its last
            // commit sits on the other mbarrier, whose next commit is this chunk's K block 1: not issued yet)
            if (kb == 0 && prev_chunk >= 0) epilogue(prev_chunk, (ci - 1) & 1, prev_tl);
        }
        prev_chunk = chunk;
        prev_tl = t - 1;
    }
    if (prev_chunk >= 0) epilogue(prev_chunk, (ci - 1) & 1, prev_tl);
    if (!alive && tid == 0) *reinterpret_cast<volatile int*>(status) = 1;  // pinned host word, read at the next sync
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_d) : "memory");
}

}  // namespace

int32_t init_lp_update_tc5(qb200_ctx* ctx) {
    QB_CUDA(ctx, cudaFuncSetAttribute(lp_update_tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LU_SMEM));
    return QB200_OK;
}

// S (rows x np float2, leading dimension lds, rows % 128 == 0): S_p <- S_p W_p for every flagged pair of the step
int32_t launch_lp_update_tc5(qb200_ctx* ctx, float2* S, int64_t lds, int64_t rows, int nb, int step, const c128* Wg,
                             const int* flags, int npairs) {
    if (rows % LU_ROWS) QB_FAIL(ctx, QB200_E_INVALID, "lp_update_tc5: rows must be a multiple of 128");
    const int nchunks = (int)(rows / LU_ROWS);
    int splits = std::max(1, ctx->sm_count / std::max(1, npairs));
    if (splits > nchunks) splits = nchunks;
    lp_update_tc5_kernel<<<dim3(splits, npairs), LU_THREADS, LU_SMEM, ctx->stream>>>(S, lds, nchunks, nb, step, Wg, flags,
                                                                                      qb_async_status(ctx));
    QB_CUDA(ctx, cudaGetLastError());  // the caller counts the launch (it may be inside a stream capture)
    return QB200_OK;
}

}  // namespace qb
