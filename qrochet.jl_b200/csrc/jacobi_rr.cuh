// Block pairing of the one-sided block-Jacobi iteration, shared by svd_jacobi.cu and jacobi_lp_tc5.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qb {

constexpr int JB = 32;   // column block width
constexpr int JP = 64;   // panel width (two blocks)

// circle-method round robin: n (even) players, round `step` in [0, n-1), pair k in [0, n/2)
__device__ __forceinline__ void rr_pair(int n, int step, int k, int& p, int& q) {
    if (step < 0) {  // consecutive blocks (2k, 2k+1): a 64-column panel addressed directly (QR panels)
        p = 2 * k;
        q = 2 * k + 1;
        return;
    }
    if (n == 2) {
        p = 0;
        q = 1;
        return;
    }
    int a, b;
    if (k == 0) {
        a = n - 1;
        b = step;
    } else {
        a = (step + k) % (n - 1);
        b = (step - k + (n - 1)) % (n - 1);
    }
    p = min(a, b);
    q = max(a, b);
}

__device__ __forceinline__ int64_t panel_col(int I, int J, int c) { return (c < JB) ? (I * JB + c) : (J * JB + c - JB); }

}  // namespace qb
