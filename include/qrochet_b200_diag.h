/*
 * libqrochet_b200_diag.so -- micro-benchmarks and hardware probes (NOT part of the product ABI).
 *
 * These measure the denominators the rooflines are quoted against (FP64 DMMA peak, mma.sync TF32/BF16 peaks,
 * tcgen05 TF32 / INT8 issue rates).  They live in their own shared library so that libqrochet_b200.so exports
 * only the drop-in boundary of include/qrochet_b200.h; the library links against libqrochet_b200.so (it runs on a
 * qb200_ctx's stream and uses its event timer).
 */
#ifndef QROCHET_B200_DIAG_H
#define QROCHET_B200_DIAG_H

#include "qrochet_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* FP64 tensor-core (DMMA m8n8k4) peak micro-benchmark: returns achieved TFLOP/s */
int32_t qb200_bench_dmma_peak(qb200_ctx* ctx, double* tflops);
/* legacy warp-level tensor path (mma.sync) peaks, dense, FP32 accumulate: tflops2 = {TF32 m16n8k8, BF16 m16n8k16};
 * the denominators for the ComplexF32 kernels */
int32_t qb200_bench_hmma_peak(qb200_ctx* ctx, double* tflops2);
/* tcgen05 / TMEM building block of the ComplexF32 path (hand-written PTX: tcgen05.alloc / mma.kind::tf32 / commit /
 * ld): out3 = {max |D - expected| of a self-checked M = 128, N = 128 / 256 product (must be 0), issue-bound TF32
 * TFLOP/s at N = 128, at N = 256} */
int32_t qb200_bench_tcgen05_tf32(qb200_ctx* ctx, double* out3);
/* the same probe on the INT8 tensor pipe (S8 x S8 -> S32 in TMEM; measured for the FP64-emulation study, profiles/r2_i8_ozaki_first_run.txt):
 * out3 = {max |D - expected| (must be 0), issue-bound TOP/s at N = 128, at N = 256} */
int32_t qb200_bench_tcgen05_i8(qb200_ctx* ctx, double* out3);
/* FP64 pipes micro-benchmark: TFLOP/s of {DMMA only, DFMA only, both issued from alternating warps} */
int32_t qb200_bench_dual_pipe(qb200_ctx* ctx, double* tflops3);
/* DMMA issue-pattern micro-benchmark: TFLOP/s for {independent accumulators, complex-multiply pattern with register
 * operands, the same with A fragments re-loaded from shared memory} x {8, 16, 32 warps per SM}; row-major [3][3] */
int32_t qb200_bench_dmma_patterns(qb200_ctx* ctx, double* tflops9);
/* 3M complex-product DMMA stream: operand sum by DADD / from a shared-memory plane / absent, at 8 and 16 warps per SM */
int32_t qb200_bench_dmma_3m(qb200_ctx* ctx, double* tflops6);
/* Jacobi update kernel (k x k matrix, microseconds per launch) against diagnostic variants that drop one ingredient each:
 * {production 3M, no operand-sum DADDs, no global stores, no cp.async refill, all three, all three and no barrier, 4M} */
int32_t qb200_bench_update_variants(qb200_ctx* ctx, int32_t k, int32_t steps, double* us7);

#ifdef __cplusplus
}
#endif
#endif
