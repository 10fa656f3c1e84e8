// FP64 tensor-core (DMMA m8n8k4) peak micro-benchmark: the roofline denominator for K1/K4/K5
// (MEASURED_PEAKS.json carries no FP64 figure).
#include "../common.cuh"
#include "../mma.cuh"

using namespace qb;

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) dmma884(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

extern "C" int32_t qb200_bench_dmma_peak(qb200_ctx* ctx, double* tflops) {
    if (!ctx || !tflops) return QB200_E_INVALID;
    Workspace ws(ctx);
    double* out = ws.get<double>(1);
    const int iters = 4096, blocks = ctx->sm_count * 4;
    dmma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(out, 64);  // warm-up
    QB_LAUNCH_CHECK(ctx);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        QB_TRY(qb200_timer_begin(ctx));
        dmma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(out, iters);
        QB_LAUNCH_CHECK(ctx);
        double ms = 0.0;
        QB_TRY(qb200_timer_end(ctx, &ms));
        double flops = (double)blocks * 8 /*warps*/ * iters * 16.0 * (2.0 * 8 * 8 * 4);
        best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return QB200_OK;
}

// Experiment: do the FP64 tensor pipe (DMMA) and the FP64 FMA pipe run concurrently?  Warps with (warp & 1) == sel
// run DMMA, the others DFMA; mode 0 = DMMA only, 1 = DFMA only, 2 = both.
__global__ void __launch_bounds__(256) dual_pipe_kernel(double* out, int iters, int mode) {
    const int warp = threadIdx.x >> 5;
    const bool do_mma = (mode == 0) || (mode == 2 && (warp & 1) == 0);
    const bool do_fma = (mode == 1) || (mode == 2 && (warp & 1) == 1);
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    if (do_mma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(acc[i], a, b);
        }
    }
    if (do_fma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                acc[i][0] = fma(a, acc[i][0], b);
                acc[i][1] = fma(b, acc[i][1], a);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;
}

extern "C" int32_t qb200_bench_dual_pipe(qb200_ctx* ctx, double* tflops3) {
    if (!ctx || !tflops3) return QB200_E_INVALID;
    Workspace ws(ctx);
    double* out = ws.get<double>(1);
    const int iters = 4096, blocks = ctx->sm_count * 4;
    for (int mode = 0; mode < 3; ++mode) {
        dual_pipe_kernel<<<blocks, 256, 0, ctx->stream>>>(out, 64, mode);
        QB_LAUNCH_CHECK(ctx);
        QB_TRY(qb200_timer_begin(ctx));
        dual_pipe_kernel<<<blocks, 256, 0, ctx->stream>>>(out, iters, mode);
        QB_LAUNCH_CHECK(ctx);
        double ms = 0.0;
        QB_TRY(qb200_timer_end(ctx, &ms));
        double warps = (double)blocks * 8;
        double mma_flops = warps * (mode == 2 ? 0.5 : (mode == 0 ? 1.0 : 0.0)) * iters * 16.0 * 512.0;
        double fma_flops = warps * (mode == 2 ? 0.5 : (mode == 1 ? 1.0 : 0.0)) * iters * 32.0 * 2.0 * 32.0;
        tflops3[mode] = (mma_flops + fma_flops) / (ms * 1e-3) / 1e12;
    }
    return QB200_OK;
}

// Experiment: which issue pattern of DMMA m8n8k4 reaches the pipe peak?  MODE 0: 16 independent accumulators,
// operands fixed (the peak benchmark); 1: the complex-multiply pattern of the update kernel (8 accumulators, each
// used twice per k-step, 4 distinct A pairs); 2: pattern 1 with the A operands re-loaded from shared memory every
// k-step (LDS.128, conflict-free pitch) and 16 different B operands from registers.
template <int MODE>
__global__ void __launch_bounds__(256) dmma_pattern_kernel(double* out, int iters) {
    __shared__ __align__(16) double2 zs[64 * 34];
    for (int i = threadIdx.x; i < 64 * 34; i += 256) zs[i] = make_double2(1.0 + i * 1e-9, 1.0 - i * 1e-9);
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double s = 0.0;
    if (MODE == 0) {
        double acc[16][2];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i][0] = acc[i][1] = 0.0;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) dmma884(acc[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i][0] + acc[i][1];
    } else {
        double cr[4][2], ci[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) cr[i][0] = cr[i][1] = ci[i][0] = ci[i][1] = 0.0;
        double2 breg[16];
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) breg[kk] = make_double2(a + kk, b - kk);
        const double2* za = zs + g;
        for (int it = 0; it < iters; it += 16) {
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                double ar[4], ai[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    if (MODE == 2) {
                        double2 v = za[(kk * 4 + t) * 34 + x * 8];
                        ar[x] = v.x;
                        ai[x] = v.y;
                    } else {
                        ar[x] = a + x;
                        ai[x] = b + x;
                    }
                }
                const double br = breg[kk].x, bi = breg[kk].y;
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    dmma884(cr[x], ar[x], br);
                    dmma884(ci[x], ar[x], bi);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                    dmma884(cr[x], -ai[x], bi);
                    dmma884(ci[x], ai[x], br);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) s += cr[i][0] + cr[i][1] + ci[i][0] + ci[i][1];
    }
    if (s == 123.456) out[0] = s;
}

// tflops[3 modes][3 occupancies: 8, 16, 32 warps per SM]
extern "C" int32_t qb200_bench_dmma_patterns(qb200_ctx* ctx, double* tflops9) {
    if (!ctx || !tflops9) return QB200_E_INVALID;
    Workspace ws(ctx);
    double* out = ws.get<double>(1);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode)
        for (int occ = 0; occ < 3; ++occ) {
            const int blocks = ctx->sm_count * (1 << occ);
            auto launch = [&](int n) {
                if (mode == 0) dmma_pattern_kernel<0><<<blocks, 256, 0, ctx->stream>>>(out, n);
                if (mode == 1) dmma_pattern_kernel<1><<<blocks, 256, 0, ctx->stream>>>(out, n);
                if (mode == 2) dmma_pattern_kernel<2><<<blocks, 256, 0, ctx->stream>>>(out, n);
            };
            launch(64);
            QB_LAUNCH_CHECK(ctx);
            QB_TRY(qb200_timer_begin(ctx));
            launch(iters);
            QB_LAUNCH_CHECK(ctx);
            double ms = 0.0;
            QB_TRY(qb200_timer_end(ctx, &ms));
            double flops = (double)blocks * 8 * iters * 16.0 * 512.0;
            tflops9[mode * 3 + occ] = flops / (ms * 1e-3) / 1e12;
        }
    return QB200_OK;
}

// Experiment (round 2): what do the operand sums of the 3M complex product cost next to the DMMA stream?  The loop of
// jacobi_update_kernel<1> (per k-step and row tile: one LDS.128, 3 DMMA) with MODE 0: the sum ar + ai formed by a DADD
// per fragment (what the kernel does), 1: the sum read from a shared-memory plane (LDS.64, no DADD), 2: no third operand
// at all (ar reused: the DMMA stream alone with its LDS.128).  3 accumulator sets x 2 row tiles, B from registers.
template <int MODE>
__global__ void __launch_bounds__(256, 2) dmma_3m_kernel(double* out, int iters) {
    __shared__ __align__(16) double2 zs[64 * 18];
    __shared__ __align__(16) double ss_plane[64 * 20];
    for (int i = threadIdx.x; i < 64 * 18; i += 256) zs[i] = make_double2(1.0 + i * 1e-9, 1.0 - i * 1e-9);
    for (int i = threadIdx.x; i < 64 * 20; i += 256) ss_plane[i] = 2.0 + i * 1e-9;
    __syncthreads();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double pp[2][2], qq[2][2], sm[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) pp[i][0] = pp[i][1] = qq[i][0] = qq[i][1] = sm[i][0] = sm[i][1] = 0.0;
    double2 breg[16];
    double bsum[16];
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
        breg[kk] = make_double2(a + kk, b - kk);
        bsum[kk] = breg[kk].x + breg[kk].y;
    }
    const double2* za = zs + g;
    const double* sa = ss_plane + g;
    for (int it = 0; it < iters; it += 16) {
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                double2 v = za[(kk * 4 + t) * 18 + x * 8];
                double s3;
                if (MODE == 0) s3 = v.x + v.y;
                else if (MODE == 1) s3 = sa[(kk * 4 + t) * 20 + x * 8];
                else s3 = v.x;
                dmma884(pp[x], v.x, breg[kk].x);
                dmma884(qq[x], v.y, breg[kk].y);
                dmma884(sm[x], s3, bsum[kk]);
            }
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 2; ++i) s += pp[i][0] + pp[i][1] + qq[i][0] + qq[i][1] + sm[i][0] + sm[i][1];
    if (s == 123.456) out[0] = s;
}

// tflops[3 modes][2 occupancies: 8, 16 warps per SM] of EXECUTED DMMA flops
extern "C" int32_t qb200_bench_dmma_3m(qb200_ctx* ctx, double* tflops6) {
    if (!ctx || !tflops6) return QB200_E_INVALID;
    Workspace ws(ctx);
    double* out = ws.get<double>(1);
    const int iters = 4096;
    for (int mode = 0; mode < 3; ++mode)
        for (int occ = 0; occ < 2; ++occ) {
            const int blocks = ctx->sm_count * (1 << occ);
            auto launch = [&](int n) {
                if (mode == 0) dmma_3m_kernel<0><<<blocks, 256, 0, ctx->stream>>>(out, n);
                if (mode == 1) dmma_3m_kernel<1><<<blocks, 256, 0, ctx->stream>>>(out, n);
                if (mode == 2) dmma_3m_kernel<2><<<blocks, 256, 0, ctx->stream>>>(out, n);
            };
            launch(64);
            QB_LAUNCH_CHECK(ctx);
            QB_TRY(qb200_timer_begin(ctx));
            launch(iters);
            QB_LAUNCH_CHECK(ctx);
            double ms = 0.0;
            QB_TRY(qb200_timer_end(ctx, &ms));
            double flops = (double)blocks * 8 * iters * 6.0 * 512.0;
            tflops6[mode * 2 + occ] = flops / (ms * 1e-3) / 1e12;
        }
    return QB200_OK;
}

// Jacobi update kernel against its own diagnostic variants (what bounds it?): us per launch at k x k for
// {production 3M, no operand-sum DADDs, no global stores, no cp.async after the prologue, all three, all three and no
// barrier, 4M production}
extern "C" int32_t qb200_bench_update_variants(qb200_ctx* ctx, int32_t k, int32_t steps, double* us7) {
    if (!ctx || !us7) return QB200_E_INVALID;
    return qb::qb_update_bench(ctx, k, steps, us7, 7);
}

// Legacy warp-level tensor path (mma.sync, SASS HMMA) peak for the ComplexF32 kernels: TF32 m16n8k8 and BF16 m16n8k16
// with FP32 accumulation, issue-bound, 8 independent accumulator tiles per warp.
__device__ __forceinline__ void mma_tf32_1688(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int KIND>
__global__ void __launch_bounds__(256) hmma_peak_kernel(float* out, int iters) {
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, 7};
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                mma_tf32_1688(acc[i], a, b);
            else
                mma_bf16_16816(acc[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
    if (s == 123.456f) out[0] = s;
}

// tflops2[0] = TF32 m16n8k8, tflops2[1] = BF16 m16n8k16 (dense, FP32 accumulate)
extern "C" int32_t qb200_bench_hmma_peak(qb200_ctx* ctx, double* tflops2) {
    if (!ctx || !tflops2) return QB200_E_INVALID;
    Workspace ws(ctx);
    float* out = ws.get<float>(1);
    const int iters = 4096, blocks = ctx->sm_count * 4;
    for (int kind = 0; kind < 2; ++kind) {
        double best = 0.0;
        for (int rep = 0; rep < 4; ++rep) {
            QB_TRY(qb200_timer_begin(ctx));
            if (kind == 0)
                hmma_peak_kernel<0><<<blocks, 256, 0, ctx->stream>>>(out, rep ? iters : 64);
            else
                hmma_peak_kernel<1><<<blocks, 256, 0, ctx->stream>>>(out, rep ? iters : 64);
            QB_LAUNCH_CHECK(ctx);
            double ms = 0.0;
            QB_TRY(qb200_timer_end(ctx, &ms));
            if (!rep) continue;
            double flops = (double)blocks * 8 * iters * 8.0 * (2.0 * 16 * 8 * (kind ? 16 : 8));
            best = std::max(best, flops / (ms * 1e-3) / 1e12);
        }
        tflops2[kind] = best;
    }
    return QB200_OK;
}
