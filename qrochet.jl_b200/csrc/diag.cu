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
