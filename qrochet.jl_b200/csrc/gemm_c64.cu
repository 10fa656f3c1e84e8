// K1 for ComplexF32: permutation-fused complex GEMM on the warp-level TF32 tensor path (mma.sync m16n8k8, SASS
// HMMA.1688.F32.TF32) with the 3xTF32 split, for sm_100a.
//
// Same contract as gemm_c128.cu (offset tables / affine strides per M, N, K and batch part of every operand; the
// index permutation of the contraction is fused into the tile loads), on float2 elements: one 8-byte cp.async per
// ComplexF32 element straight from its permuted location.
//
// Arithmetic: every FP32 operand x is split on the fly into two TF32 numbers, x = hi + lo with hi = tf32(x) and
// lo = tf32(x - hi); a real product block is accumulated in FP32 as lo*hi + hi*lo + hi*hi (small terms first), which
// recovers FP32-level accuracy (relative error ~1e-6 on a K = 4096 dot product instead of ~1e-3 for plain TF32) --
// the 1e-5 ComplexF32 tolerance of the north star needs it.  A complex 16x8x8 step is 4 real products = 12 HMMA.
//
// CTA tile 128 x 64 x 16 complex, 512 threads = 16 warps as 4(M) x 4(N), warp tile 32 x 16 complex = 2 x 2 HMMA
// tiles x (re, im) = 32 accumulator registers.  4-stage cp.async pipeline; shared tiles are k-major with pitch + 4
// (in float2) so that the 8-byte fragment loads of a half warp (4 k x 4 m) hit 16 distinct 8-byte bank pairs.
#include "gemm_c128.cuh"
#include "mma.cuh"

#include <algorithm>

namespace qb {

namespace {

constexpr int FBM = 128, FBN = 64, FBK = 16, FSTAGES = 4;
constexpr int FPA = FBM + 4, FPB = FBN + 4;
constexpr int F_THREADS = 512;
constexpr size_t GEMM32_SMEM = (size_t)FSTAGES * FBK * (FPA + FPB) * sizeof(float2);

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src, bool valid) {
    uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem_dst);
    int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gmem_src), "r"(sz) : "memory");
}

// x = hi + lo with hi = x rounded to TF32 (10-bit mantissa, round to nearest done with integer arithmetic: add half a
// TF32 ulp to the bit pattern, clear the 13 low bits) and lo = x - hi (exact in FP32; the tensor core reads its top
// 10 mantissa bits).  IADD + LOP3 + FADD run at full rate; cvt.rna.tf32 is a quarter-rate conversion-pipe
// instruction and two of them per value made the kernel conversion-bound (measured: 45 -> see profiles/).
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

// D(16x8) += A(16x8, row) * B(8x8, col), TF32 inputs, FP32 accumulate.  lane = 4 g + t:
//   a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]; b0 = B[t][g], b1 = B[t+4][g];
//   c0 = C[g][2t], c1 = C[g][2t+1], c2 = C[g+8][2t], c3 = C[g+8][2t+1]
__device__ __forceinline__ void hmma1688(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// c += a * b with the 3xTF32 split (a = ah + al, b = bh + bl; the al*bl term is below FP32 resolution)
__device__ __forceinline__ void mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
    hmma1688(c, al, bh);
    hmma1688(c, ah, bl);
    hmma1688(c, ah, bh);
}

struct Frag32A {  // one 16 x 8 complex A fragment split into TF32 pairs
    uint32_t rh[4], rl[4], ih[4], il[4];
};
struct Frag32B {  // one 8 x 8 complex B fragment; nih / nil hold -imag (for the real part of the product)
    uint32_t rh[2], rl[2], ih[2], il[2], nih[2], nil[2];
};

// AFFINE = true: none of the M / N / K parts needs an offset table (plain strided matrices, or mode groups that
// merged into one stride).  The loader then walks K with one pointer increment per element and stage instead of
// re-deriving every address (64-bit multiply-adds, table-or-stride branches: ~100 instructions per element, which
// made the kernel issue-bound at 48 % of the HMMA pipe -- the TF32 pipe is 4x faster per flop than DMMA, so the
// loader overhead the ComplexF64 kernel hides shows here).
template <bool AFFINE>
__global__ void __launch_bounds__(F_THREADS, 1) gemm_c64_kernel(const GemmArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* As = reinterpret_cast<float2*>(smem_raw);  // [FSTAGES][FBK][FPA]
    float2* Bs = As + (size_t)FSTAGES * FBK * FPA;     // [FSTAGES][FBK][FPB]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int g = lane >> 2, t = lane & 3;
    int tile_m, tile_n;
    {  // grouped rasterisation, as in the ComplexF64 kernel
        const int gm = gridDim.x, gn = gridDim.y, GROUP = 8;
        const int id = blockIdx.x + gm * blockIdx.y;
        const int per_group = GROUP * gn;
        const int first_m = (id / per_group) * GROUP;
        const int gsz = min(gm - first_m, GROUP);
        tile_m = first_m + (id % per_group) % gsz;
        tile_n = (id % per_group) / gsz;
    }
    const int m0 = tile_m * FBM, n0 = tile_n * FBN;
    const int z = (p.ksplit > 1) ? 0 : blockIdx.z;
    const int split = (p.ksplit > 1) ? blockIdx.z : 0;

    const float2* __restrict__ A = reinterpret_cast<const float2*>(p.A) + p.ab.at(z);
    const float2* __restrict__ B = reinterpret_cast<const float2*>(p.B) + p.bb.at(z);
    float2* __restrict__ C = reinterpret_cast<float2*>(p.C) + p.cb.at(z);

    constexpr int A_PER = FBM * FBK / F_THREADS, B_PER = FBN * FBK / F_THREADS;
    int a_ml[A_PER], a_kl[A_PER];
    int64_t a_moff[A_PER];
    bool a_ok[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int e = tid + F_THREADS * i;
        if (p.a_kfast) {
            a_kl[i] = e % FBK;
            a_ml[i] = e / FBK;
        } else {
            a_ml[i] = e % FBM;
            a_kl[i] = e / FBM;
        }
        a_ok[i] = (m0 + a_ml[i]) < p.M;
        a_moff[i] = a_ok[i] ? p.am.at(m0 + a_ml[i]) : 0;
    }
    int b_nl[B_PER], b_kl[B_PER];
    int64_t b_noff[B_PER];
    bool b_ok[B_PER];
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
        int e = tid + F_THREADS * i;
        if (p.b_kfast) {
            b_kl[i] = e % FBK;
            b_nl[i] = e / FBK;
        } else {
            b_nl[i] = e % FBN;
            b_kl[i] = e / FBN;
        }
        b_ok[i] = (n0 + b_nl[i]) < p.N;
        b_noff[i] = b_ok[i] ? p.bn.at(n0 + b_nl[i]) : 0;
    }

    const int KT_all = (p.K + FBK - 1) / FBK;
    const int kt_per = (KT_all + p.ksplit - 1) / p.ksplit;
    const int kt0 = split * kt_per;
    const int KT = max(0, min(KT_all, kt0 + kt_per) - kt0);

    // affine walk: per-element source pointers advanced by one K tile per call (tiles are loaded in order)
    const float2* a_ptr[A_PER];
    const float2* b_ptr[B_PER];
    int a_so[A_PER], b_so[B_PER];  // shared-memory offsets inside a stage
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        a_so[i] = a_kl[i] * FPA + a_ml[i];
        a_ptr[i] = AFFINE ? A + a_moff[i] + (int64_t)(kt0 * FBK + a_kl[i]) * p.ak.stride : A;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
        b_so[i] = b_kl[i] * FPB + b_nl[i];
        b_ptr[i] = AFFINE ? B + b_noff[i] + (int64_t)(kt0 * FBK + b_kl[i]) * p.bk.stride : B;
    }
    const int64_t a_step = (int64_t)FBK * p.ak.stride, b_step = (int64_t)FBK * p.bk.stride;

    auto load_tile = [&](int kt, int s) {
        float2* as = As + (size_t)s * FBK * FPA;
        float2* bs = Bs + (size_t)s * FBK * FPB;
        if constexpr (AFFINE) {
            const int kbase = (kt0 + kt) * FBK;
#pragma unroll
            for (int i = 0; i < A_PER; ++i) {
                const bool ok = a_ok[i] && (kbase + a_kl[i]) < p.K;
                cp_async8(as + a_so[i], ok ? a_ptr[i] : reinterpret_cast<const float2*>(p.A), ok);
                a_ptr[i] += a_step;
            }
#pragma unroll
            for (int i = 0; i < B_PER; ++i) {
                const bool ok = b_ok[i] && (kbase + b_kl[i]) < p.K;
                cp_async8(bs + b_so[i], ok ? b_ptr[i] : reinterpret_cast<const float2*>(p.B), ok);
                b_ptr[i] += b_step;
            }
        } else {
#pragma unroll
            for (int i = 0; i < A_PER; ++i) {
                int kg = (kt0 + kt) * FBK + a_kl[i];
                bool ok = a_ok[i] && kg < p.K;
                const float2* src = ok ? (A + a_moff[i] + p.ak.at(kg)) : reinterpret_cast<const float2*>(p.A);
                cp_async8(as + a_so[i], src, ok);
            }
#pragma unroll
            for (int i = 0; i < B_PER; ++i) {
                int kg = (kt0 + kt) * FBK + b_kl[i];
                bool ok = b_ok[i] && kg < p.K;
                const float2* src = ok ? (B + b_noff[i] + p.bk.at(kg)) : reinterpret_cast<const float2*>(p.B);
                cp_async8(bs + b_so[i], src, ok);
            }
        }
    };

    float accr[2][2][4], acci[2][2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 4; ++h) accr[i][j][h] = acci[i][j][h] = 0.f;

#pragma unroll
    for (int s = 0; s < FSTAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    const float sa = p.conjA ? -1.f : 1.f, sb = p.conjB ? -1.f : 1.f;

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<FSTAGES - 2>();
        __syncthreads();
        {
            int nk = kt + FSTAGES - 1;
            if (nk < KT) load_tile(nk, nk % FSTAGES);
            cp_async_commit();
        }
        const float2* as = As + (size_t)(kt % FSTAGES) * FBK * FPA + wm * 32 + g;
        const float2* bs = Bs + (size_t)(kt % FSTAGES) * FBK * FPB + wn * 16 + g;
#pragma unroll
        for (int kk = 0; kk < FBK / 8; ++kk) {
            Frag32B fb[2];
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {  // b0 = B[t][g], b1 = B[t+4][g]
                    float2 v = bs[(kk * 8 + t + 4 * e) * FPB + j * 8];
                    v.y *= sb;
                    split_tf32(v.x, fb[j].rh[e], fb[j].rl[e]);
                    split_tf32(v.y, fb[j].ih[e], fb[j].il[e]);
                    fb[j].nih[e] = fb[j].ih[e] ^ 0x80000000u;
                    fb[j].nil[e] = fb[j].il[e] ^ 0x80000000u;
                }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                Frag32A fa;
#pragma unroll
                for (int e = 0; e < 4; ++e) {  // a0 = A[g][t], a1 = A[g+8][t], a2 = A[g][t+4], a3 = A[g+8][t+4]
                    float2 v = as[(kk * 8 + t + 4 * (e >> 1)) * FPA + i * 16 + 8 * (e & 1)];
                    v.y *= sa;
                    split_tf32(v.x, fa.rh[e], fa.rl[e]);
                    split_tf32(v.y, fa.ih[e], fa.il[e]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    mma3(accr[i][j], fa.rh, fa.rl, fb[j].rh, fb[j].rl);    // re += ar br
                    mma3(accr[i][j], fa.ih, fa.il, fb[j].nih, fb[j].nil);  // re -= ai bi
                    mma3(acci[i][j], fa.rh, fa.rl, fb[j].ih, fb[j].il);    // im += ar bi
                    mma3(acci[i][j], fa.ih, fa.il, fb[j].rh, fb[j].rl);    // im += ai br
                }
            }
        }
    }
    cp_async_wait<0>();

    if (p.ksplit > 1) {
        float2* part = reinterpret_cast<float2*>(p.partial) + (size_t)split * p.M * p.N;
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    int m = m0 + wm * 32 + i * 16 + g + 8 * (h >> 1);
                    int n = n0 + wn * 16 + j * 8 + 2 * t + (h & 1);
                    if (m < p.M && n < p.N) part[m + (size_t)p.M * n] = make_float2(accr[i][j][h], acci[i][j][h]);
                }
        return;
    }
    const float alr = (float)p.alpha.x, ali = (float)p.alpha.y, ber = (float)p.beta.x, bei = (float)p.beta.y;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
            int m = m0 + wm * 32 + i * 16 + g + 8 * hr;
            if (m >= p.M) continue;
            int64_t mo = p.cm.at(m);
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {
                    int n = n0 + wn * 16 + j * 8 + 2 * t + hc;
                    if (n >= p.N) continue;
                    float2* dst = C + mo + p.cn.at(n);
                    float vr = accr[i][j][2 * hr + hc], vi = acci[i][j][2 * hr + hc];
                    float2 o = make_float2(alr * vr - ali * vi, alr * vi + ali * vr);
                    if (!p.beta_zero) {
                        float2 old = *dst;
                        o.x += ber * old.x - bei * old.y;
                        o.y += ber * old.y + bei * old.x;
                    }
                    *dst = o;
                }
        }
}

// sums the split-K partials in a fixed order (deterministic, FP64 running sum) and applies alpha / beta
__global__ void splitk_reduce_c64_kernel(const GemmArgs p) {
    int64_t total = (int64_t)p.M * p.N;
    const float2* part = reinterpret_cast<const float2*>(p.partial);
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int m = (int)(idx % p.M), n = (int)(idx / p.M);
        double vr = 0.0, vi = 0.0;
        for (int s = 0; s < p.ksplit; ++s) {
            float2 v = part[(size_t)s * total + idx];
            vr += v.x;
            vi += v.y;
        }
        float2* dst = reinterpret_cast<float2*>(p.C) + p.cm.at(m) + p.cn.at(n) + p.cb.at(0);
        double orr = p.alpha.x * vr - p.alpha.y * vi, oi = p.alpha.x * vi + p.alpha.y * vr;
        if (!p.beta_zero) {
            float2 old = *dst;
            orr += p.beta.x * old.x - p.beta.y * old.y;
            oi += p.beta.x * old.y + p.beta.y * old.x;
        }
        *dst = make_float2((float)orr, (float)oi);
    }
}

}  // namespace

int32_t init_gemm_c64(qb200_ctx* ctx) {
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM32_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM32_SMEM));
    return QB200_OK;
}

// `args` as built for the ComplexF64 kernel; A / B / C / partial point at float2 data
int32_t launch_gemm_c64(qb200_ctx* ctx, const GemmArgs& args_in) {
    if (args_in.M <= 0 || args_in.N <= 0 || args_in.batch <= 0) return QB200_OK;
    GemmArgs args = args_in;
    if (args.N > args.M) {  // C^T = B^T A^T: the larger free dimension goes on the 128-wide side
        std::swap(args.A, args.B);
        std::swap(args.am, args.bn);
        std::swap(args.ak, args.bk);
        std::swap(args.ab, args.bb);
        std::swap(args.cm, args.cn);
        std::swap(args.M, args.N);
        std::swap(args.conjA, args.conjB);
        std::swap(args.a_kfast, args.b_kfast);
    }
    args.ksplit = 1;
    args.partial = nullptr;
    args.acc_init = 0;
    Workspace ws(ctx);
    {
        int64_t tiles = (int64_t)((args.M + FBM - 1) / FBM) * ((args.N + FBN - 1) / FBN);
        int KT = (args.K + FBK - 1) / FBK;
        int want = 1;
        if (args.batch == 1 && tiles * 2 <= ctx->sm_count && KT >= 32)
            want = (int)std::min<int64_t>(ctx->sm_count / tiles, KT / 8);
        // the tcgen05 kernel (gemm_c64_tc5.cu) takes every product except the few-tiles / long-K ones that need
        // split-K to fill the machine, which stay on the mma.sync kernel below
        if (want <= 1 && gemm_c64_tc5_enabled()) return launch_gemm_c64_tc5(ctx, args);
        if (args.batch == 1 && tiles * 2 <= ctx->sm_count && KT >= 32) {
            if (want > 1) {
                args.partial = reinterpret_cast<c128*>(ws.get<float2>((size_t)want * args.M * args.N));
                if (!args.partial) QB_FAIL(ctx, QB200_E_CUDA, "gemm_c64: split-K workspace allocation failed");
                args.ksplit = want;
            }
        }
    }
    dim3 grid((args.M + FBM - 1) / FBM, (args.N + FBN - 1) / FBN, args.ksplit > 1 ? args.ksplit : args.batch);
    if (grid.y > 65535 || grid.z > 65535) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "gemm_c64 grid too large");
    const bool affine = !args.am.tab && !args.ak.tab && !args.bk.tab && !args.bn.tab;
    if (affine)
        gemm_c64_kernel<true><<<grid, F_THREADS, GEMM32_SMEM, ctx->stream>>>(args);
    else
        gemm_c64_kernel<false><<<grid, F_THREADS, GEMM32_SMEM, ctx->stream>>>(args);
    QB_LAUNCH_CHECK(ctx);
    if (args.ksplit > 1) {
        int64_t total = (int64_t)args.M * args.N;
        unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 8);
        splitk_reduce_c64_kernel<<<blocks, 256, 0, ctx->stream>>>(args);
        QB_LAUNCH_CHECK(ctx);
    }
    return QB200_OK;
}

}  // namespace qb
