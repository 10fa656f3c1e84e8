// tcgen05 / TMEM primitives shared by the ComplexF32 kernels (gemm_c64_tc5.cu, jacobi_lp_tc5.cu): canonical K-major
// no-swizzle UMMA operand layout (8-row x 16-byte core matrices, SBO = 128 B between 8-row groups, LBO = 2048 B between
// 4-wide K chunks of a 128-row operand), the kind::tf32 MMA, mbarrier wait with a timeout, the round-to-nearest TF32
// hi / lo split and the 32-column TMEM load.  Hand-written PTX, validated by diag/tc5_probe.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qb {
namespace tc5 {

constexpr uint32_t T_SBO = 128, T_LBO = 128 * 16;             // both operand tiles have 128 rows
constexpr uint32_t T_OPER_BYTES = 16 * T_LBO;                 // 16 K chunks of 4 = 64 real K: 32 KB per operand

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((T_LBO >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((T_SBO >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version: Blackwell
    return d;
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

__device__ __forceinline__ void split4(const float (&x)[4], float4& hi, float4& lo) {
    float h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = __uint_as_float((__float_as_uint(x[j]) + 0x1000u) & 0xffffe000u);  // round to nearest TF32
        l[j] = x[j] - h[j];                                                      // exact; the MMA reads its top bits
    }
    hi = make_float4(h[0], h[1], h[2], h[3]);
    lo = make_float4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float4 neg4(float4 v) { return make_float4(-v.x, -v.y, -v.z, -v.w); }

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

}  // namespace tc5
}  // namespace qb
