// FP64 tensor-core (DMMA m8n8k4) peak micro-benchmark: the roofline denominator for K1/K4/K5
// (MEASURED_PEAKS.json carries no FP64 figure).
#include "common.cuh"
#include "mma.cuh"

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
