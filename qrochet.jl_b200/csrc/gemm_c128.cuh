// K1: permutation-fused ComplexF64 GEMM on the FP64 tensor pipe (DMMA m8n8k4) for sm_100a.
//
// C[cm(i) + cn(j) + cb(z)] = alpha * sum_k opA(A[am(i) + ak(k) + ab(z)]) * opB(B[bk(k) + bn(j) + bb(z)]) + beta * C
//
// The element offset of every operand is split into an M/N/K (and batch) part, each either affine
// (index * stride) or an int64 offset table built on the device from the tensor's mode list.  That is
// how the index permutation of a tensor contraction is fused into the tile loads: the loader issues one
// 16-byte cp.async (= one ComplexF64) per element straight from the permuted location into shared
// memory, so there is no TTGT copy of the operands and no output permute.
//
// CTA tile 128 x 64 x 8 complex, 512 threads = 16 warps as 4(M) x 4(N), warp tile 32 x 16 complex =
// 4 x 2 DMMA tiles x (re, im): 64 accumulator registers.  4 real DMMAs per complex 8x8x4 step.
// 4-stage cp.async pipeline; shared tiles are k-major with pitch +2 so that the 16-byte fragment loads
// of a quarter warp hit 8 distinct 16-byte bank groups.
#pragma once
#include "common.cuh"

namespace qb {

constexpr int BM = 128, BN = 64, BK = 8, STAGES = 4;
constexpr int PA = BM + 2, PB = BN + 2;
constexpr int GEMM_THREADS = 512;  // 16 warps as 4(M) x 4(N), warp tile 32 x 16: 4 warps per SMSP hide LDS / barrier stalls
constexpr size_t GEMM_SMEM = (size_t)STAGES * BK * (PA + PB) * sizeof(c128);
constexpr size_t GEMM_SMEM_HALF = (size_t)STAGES * BK * (64 + 2 + PB) * sizeof(c128);  // 64-row tile variant

struct Operand {
    const int64_t* tab;  // offset table or null
    int64_t stride;      // used when tab == null
    __device__ __forceinline__ int64_t at(int64_t i) const { return tab ? tab[i] : i * stride; }
};

struct GemmArgs {
    const c128* A;
    const c128* B;
    c128* C;
    Operand am, ak, bk, bn, cm, cn, ab, bb, cb;
    int M, N, K, batch;
    int conjA, conjB;
    int a_kfast, b_kfast;  // loader mapping: consecutive threads walk k (1) or m/n (0)
    c128 alpha, beta;
    int beta_zero;
    int ksplit;      // set by launch_gemm
    c128* partial;   // split-K partial sums (ksplit x M x N), set by launch_gemm
    int acc_init;    // C += (+-1) A B folded into the accumulators, set by launch_gemm
};

struct ModeList {
    int n;
    int64_t ext[32];
    int64_t stride[32];
};

int32_t launch_gemm(qb200_ctx* ctx, const GemmArgs& args);
int32_t init_gemm(qb200_ctx* ctx);  // kernel attributes, once per process (called by qb200_create)
// ComplexF32 twin (gemm_c64.cu): same GemmArgs, A / B / C / partial point at float2 data
int32_t launch_gemm_c64(qb200_ctx* ctx, const GemmArgs& args);
int32_t init_gemm_c64(qb200_ctx* ctx);
// ComplexF32 on tcgen05 / TMEM (gemm_c64_tc5.cu); opt-in with QB200_C64_TCGEN05=1
int32_t launch_gemm_c64_tc5(qb200_ctx* ctx, const GemmArgs& args);
int32_t init_gemm_c64_tc5(qb200_ctx* ctx);
bool gemm_c64_tc5_enabled();
// offsets[idx] = sum_j coord_j(idx) * stride[j], first mode fastest
int32_t build_offsets(qb200_ctx* ctx, const ModeList& ml, int64_t total, int64_t* out);

}  // namespace qb
