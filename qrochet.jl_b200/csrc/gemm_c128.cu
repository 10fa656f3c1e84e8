// K1 kernel: see gemm_c128.cuh for the design.
#include "gemm_c128.cuh"
#include "mma.cuh"

#include <algorithm>
#include <cstdlib>

namespace qb {

// M3 = 1: 3M complex product (Gauss).  Per complex 8x8x4 tile P += Ar Br, Q += Ai Bi, S += (Ar + Ai)(Br + Bi) and
// Cr = P - Q, Ci = S - P - Q at the end: 3 DMMA + operand sums (DADD) instead of 4 DMMA.  The component-wise error
// bound becomes eps (|Ar| + |Ai|)(|Br| + |Bi|) -- norm-wise the same class as the 4M product.  accr/acci/accs hold
// P/Q/S in that mode.
// HALF = 1: 64-row tile, 256 threads (8 warps as 2 x 4), 68 KB: TWO CTAs per SM.  For short K (a few k tiles per CTA) the
// prologue (offset tables -> operands) and the epilogue (offset tables -> stores) of one CTA overlap the DMMA stream of
// the other; with the 128-row tile's single resident CTA they are exposed (measured on the N = 64, K = 64 node of the
// sliced benchmark network: 21 TFLOP/s).
template <int M3, int HALF>
__global__ void __launch_bounds__(HALF ? 256 : GEMM_THREADS, HALF ? 2 : 1) gemm_c128_kernel(const GemmArgs p) {
    constexpr int TBM = HALF ? 64 : BM, TPA = TBM + 2, TTHREADS = TBM * 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* As = reinterpret_cast<c128*>(smem_raw);  // [STAGES][BK][TPA]
    c128* Bs = As + (size_t)STAGES * BK * TPA;      // [STAGES][BK][PB]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % (TBM / 32), wn = warp / (TBM / 32);  // wn in 0..3: 16 columns each
    const int g = lane >> 2, t = lane & 3;
    // grouped rasterisation: consecutive CTAs walk 8 M-tiles x all N-tiles, so the tiles resident at any time share
    // A and B panels that fit in L2 (a plain M-fastest order streams all of A from HBM once per wave)
    int tile_m = blockIdx.x, tile_n = blockIdx.y;
    {
        const int gm = gridDim.x, gn = gridDim.y, GROUP = 8;
        const int id = blockIdx.x + gm * blockIdx.y;
        const int per_group = GROUP * gn;
        const int first_m = (id / per_group) * GROUP;
        const int gsz = min(gm - first_m, GROUP);
        tile_m = first_m + (id % per_group) % gsz;
        tile_n = (id % per_group) / gsz;
    }
    const int m0 = tile_m * TBM, n0 = tile_n * BN;
    // blockIdx.z is the batch entry, or the K split when split-K is on (batch == 1 then)
    const int z = (p.ksplit > 1) ? 0 : blockIdx.z;
    const int split = (p.ksplit > 1) ? blockIdx.z : 0;

    const c128* __restrict__ A = p.A + p.ab.at(z);
    const c128* __restrict__ B = p.B + p.bb.at(z);
    c128* __restrict__ C = p.C + p.cb.at(z);

    // ---- loader coordinates (fixed for the whole K loop) ----
    constexpr int A_PER = TBM * BK / TTHREADS, B_PER = BN * BK / TTHREADS;
    int a_ml[A_PER], a_kl[A_PER];
    int64_t a_moff[A_PER];
    bool a_ok[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        int e = tid + TTHREADS * i;
        if (p.a_kfast) {
            a_kl[i] = e % BK;
            a_ml[i] = e / BK;
        } else {
            a_ml[i] = e % TBM;
            a_kl[i] = e / TBM;
        }
        a_ok[i] = (m0 + a_ml[i]) < p.M;
        a_moff[i] = a_ok[i] ? p.am.at(m0 + a_ml[i]) : 0;
    }
    int b_nl[B_PER], b_kl[B_PER];
    int64_t b_noff[B_PER];
    bool b_ok[B_PER];
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
        int e = tid + TTHREADS * i;
        if (p.b_kfast) {
            b_kl[i] = e % BK;
            b_nl[i] = e / BK;
        } else {
            b_nl[i] = e % BN;
            b_kl[i] = e / BN;
        }
        b_ok[i] = (n0 + b_nl[i]) < p.N;
        b_noff[i] = b_ok[i] ? p.bn.at(n0 + b_nl[i]) : 0;
    }

    const int KT_all = (p.K + BK - 1) / BK;
    const int kt_per = (KT_all + p.ksplit - 1) / p.ksplit;
    const int kt0 = split * kt_per;
    const int KT = max(0, min(KT_all, kt0 + kt_per) - kt0);  // k tiles handled by this CTA

    auto load_tile = [&](int kt, int s) {
        c128* as = As + (size_t)s * BK * TPA;
        c128* bs = Bs + (size_t)s * BK * PB;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            int kg = (kt0 + kt) * BK + a_kl[i];
            bool ok = a_ok[i] && kg < p.K;
            const c128* src = ok ? (A + a_moff[i] + p.ak.at(kg)) : p.A;
            cp_async16(as + a_kl[i] * TPA + a_ml[i], src, ok);
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            int kg = (kt0 + kt) * BK + b_kl[i];
            bool ok = b_ok[i] && kg < p.K;
            const c128* src = ok ? (B + b_noff[i] + p.bk.at(kg)) : p.B;
            cp_async16(bs + b_kl[i] * PB + b_nl[i], src, ok);
        }
    };

    constexpr int NJ = 2;  // 8-column tiles per warp
    double accr[4][NJ][2], acci[4][NJ][2], accs[M3 ? 4 : 1][NJ][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            accr[i][j][0] = accr[i][j][1] = 0.0;
            acci[i][j][0] = acci[i][j][1] = 0.0;
            if constexpr (M3) accs[i][j][0] = accs[i][j][1] = 0.0;
        }
    if (p.acc_init) {
        // C += (+-1) A B: start the accumulators from +-C so that the epilogue is a pure store (the loads overlap
        // the pipeline prologue instead of serialising behind the main loop)
        const double sgn = p.alpha.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + wm * 32 + i * 8 + g;
            if (m >= p.M) continue;
            int64_t mo = p.cm.at(m);
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int n = n0 + wn * 16 + j * 8 + 2 * t + h;
                    if (n >= p.N) continue;
                    c128 v = C[mo + p.cn.at(n)];
                    accr[i][j][h] = sgn * v.x;
                    if constexpr (M3) {  // P = c_r, Q = 0, S = c_r + c_i
                        accs[i][j][h] = sgn * (v.x + v.y);
                    } else {
                        acci[i][j][h] = sgn * v.y;
                    }
                }
        }
    }

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }

    const int sgnA = p.conjA ? 0x80000000 : 0, sgnB = p.conjB ? 0x80000000 : 0;

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            int nk = kt + STAGES - 1;
            if (nk < KT) load_tile(nk, nk % STAGES);
            cp_async_commit();
        }
        const c128* as = As + (size_t)(kt % STAGES) * BK * TPA + wm * 32 + g;
        const c128* bs = Bs + (size_t)(kt % STAGES) * BK * PB + wn * 16 + g;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            if constexpr (M3) {
                double br[NJ], bi[NJ], bsum[NJ];
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    c128 v = bs[(kk * 4 + t) * PB + j * 8];
                    br[j] = v.x;
                    bi[j] = flip_sign(v.y, sgnB);
                    bsum[j] = br[j] + bi[j];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    c128 v = as[(kk * 4 + t) * TPA + i * 8];
                    const double ar = v.x, ai = flip_sign(v.y, sgnA), asum = ar + ai;
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        dmma884(accr[i][j], ar, br[j]);
                        dmma884(acci[i][j], ai, bi[j]);
                        dmma884(accs[i][j], asum, bsum[j]);
                    }
                }
            } else {
            double ar[4], ai[4], nai[4], br[NJ], bi[NJ];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                c128 v = as[(kk * 4 + t) * TPA + i * 8];
                ar[i] = v.x;
                ai[i] = flip_sign(v.y, sgnA);
                nai[i] = flip_sign(ai[i], 0x80000000);
            }
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                c128 v = bs[(kk * 4 + t) * PB + j * 8];
                br[j] = v.x;
                bi[j] = flip_sign(v.y, sgnB);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    dmma884(accr[i][j], ar[i], br[j]);
                    dmma884(acci[i][j], ar[i], bi[j]);
                    dmma884(accr[i][j], nai[i], bi[j]);
                    dmma884(acci[i][j], ai[i], br[j]);
                }
            }
        }
    }
    cp_async_wait<0>();
    if constexpr (M3) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const double P = accr[i][j][h], Q = acci[i][j][h];
                    accr[i][j][h] = P - Q;
                    acci[i][j][h] = accs[i][j][h] - P - Q;
                }
    }

    if (p.ksplit > 1) {
        // split-K: raw partial sums, dense M x N per split; splitk_reduce_kernel finishes the job
        c128* part = p.partial + (size_t)split * p.M * p.N;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int m = m0 + wm * 32 + i * 8 + g;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < NJ; ++j)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    int n = n0 + wn * 16 + j * 8 + 2 * t + h;
                    if (n < p.N) part[m + (size_t)p.M * n] = make_double2(accr[i][j][h], acci[i][j][h]);
                }
        }
        return;
    }
    // ---- epilogue: C = alpha * acc + beta * C, written through the C offset tables ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + wm * 32 + i * 8 + g;
        if (m >= p.M) continue;
        int64_t mo = p.cm.at(m);
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                int n = n0 + wn * 16 + j * 8 + 2 * t + h;
                if (n >= p.N) continue;
                c128* dst = C + mo + p.cn.at(n);
                double vr = accr[i][j][h], vi = acci[i][j][h];
                c128 o;
                o.x = p.alpha.x * vr - p.alpha.y * vi;
                o.y = p.alpha.x * vi + p.alpha.y * vr;
                if (!p.beta_zero && !p.acc_init) {
                    c128 old = *dst;
                    o.x += p.beta.x * old.x - p.beta.y * old.y;
                    o.y += p.beta.x * old.y + p.beta.y * old.x;
                }
                *dst = o;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Thin variant for the shapes a sliced circuit network is made of: a huge free dimension against a gate-sized operand
// (N <= 16 after orientation, K of 4 .. 128).  These contractions are HBM-bound streams (16 .. 64 flop per byte of the big
// operand), and the 128 x 64 tile kernel above runs them latency-bound: one 512-thread CTA per SM (128 registers), each
// with three dependent global round trips (offset table -> operand -> offset table -> store) and a tile that is 7/8
// padding.  Here a CTA is 128 threads on a 128 x (8 NT) tile with 2 cp.async stages (35 KB): a dozen CTAs per SM keep
// the memory system busy.  Same GemmArgs contract, fragment layout and (4M) complex product; no split-K.
constexpr int TH_THREADS = 128, TH_STAGES = 2;
template <int NT>
struct ThinCfg {
    static constexpr int BNT = 8 * NT, PBT = BNT + 2;
    static constexpr size_t SMEM = (size_t)TH_STAGES * BK * (PA + PBT) * sizeof(c128);
};

template <int NT>
__global__ void __launch_bounds__(TH_THREADS, NT == 1 ? 4 : 3) gemm_c128_thin_kernel(const GemmArgs p, int ntiles) {
    constexpr int BNT = ThinCfg<NT>::BNT, PBT = ThinCfg<NT>::PBT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    c128* As = reinterpret_cast<c128*>(smem_raw);      // [TH_STAGES][BK][PA]
    c128* Bs = As + (size_t)TH_STAGES * BK * PA;       // [TH_STAGES][BK][PBT]
    const int tid = threadIdx.x, lane = tid & 31, wm = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    // N tiles fastest: the CTAs that share an A tile are resident together (A comes from HBM once, from L2 afterwards)
    const int m0 = (blockIdx.x / ntiles) * BM, n0 = (blockIdx.x % ntiles) * BNT, z = blockIdx.z;
    const c128* __restrict__ A = p.A + p.ab.at(z);
    const c128* __restrict__ B = p.B + p.bb.at(z);
    c128* __restrict__ C = p.C + p.cb.at(z);

    constexpr int A_PER = BM * BK / TH_THREADS;  // 8
    int a_ml[A_PER], a_kl[A_PER];
    int64_t a_moff[A_PER];
    bool a_ok[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        const int e = tid + TH_THREADS * i;
        if (p.a_kfast) {
            a_kl[i] = e % BK;
            a_ml[i] = e / BK;
        } else {
            a_ml[i] = e % BM;
            a_kl[i] = e / BM;
        }
        a_ok[i] = (m0 + a_ml[i]) < p.M;
        a_moff[i] = a_ok[i] ? p.am.at(m0 + a_ml[i]) : 0;
    }
    // B tile: BK x BNT <= 128 elements, at most one per thread
    const bool b_item = tid < BK * BNT;
    const int b_kl = p.b_kfast ? tid % BK : tid / BNT, b_nl = p.b_kfast ? tid / BK : tid % BNT;
    const bool b_ok = b_item && (n0 + b_nl) < p.N;
    const int64_t b_noff = b_ok ? p.bn.at(n0 + b_nl) : 0;

    const int KT = (p.K + BK - 1) / BK;
    auto load_tile = [&](int kt, int s) {
        c128* as = As + (size_t)s * BK * PA;
        c128* bs = Bs + (size_t)s * BK * PBT;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int kg = kt * BK + a_kl[i];
            const bool ok = a_ok[i] && kg < p.K;
            cp_async16(as + a_kl[i] * PA + a_ml[i], ok ? (A + a_moff[i] + p.ak.at(kg)) : p.A, ok);
        }
        if (b_item) {
            const int kg = kt * BK + b_kl;
            const bool ok = b_ok && kg < p.K;
            cp_async16(bs + b_kl * PBT + b_nl, ok ? (B + b_noff + p.bk.at(kg)) : p.B, ok);
        }
    };
    double accr[4][NT][2], acci[4][NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) accr[i][j][0] = accr[i][j][1] = acci[i][j][0] = acci[i][j][1] = 0.0;

    load_tile(0, 0);
    cp_async_commit();
    const int sgnA = p.conjA ? 0x80000000 : 0, sgnB = p.conjB ? 0x80000000 : 0;
    for (int kt = 0; kt < KT; ++kt) {
        if (kt + 1 < KT) load_tile(kt + 1, (kt + 1) & 1);  // the other stage was released by the barrier that ended tile kt-1
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const c128* as = As + (size_t)(kt & 1) * BK * PA + wm * 32 + g;
        const c128* bs = Bs + (size_t)(kt & 1) * BK * PBT + g;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double br[NT], bi[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const c128 v = bs[(kk * 4 + t) * PBT + j * 8];
                br[j] = v.x;
                bi[j] = flip_sign(v.y, sgnB);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const c128 v = as[(kk * 4 + t) * PA + i * 8];
                const double ar = v.x, ai = flip_sign(v.y, sgnA), nai = flip_sign(ai, 0x80000000);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    dmma884(accr[i][j], ar, br[j]);
                    dmma884(acci[i][j], ar, bi[j]);
                    dmma884(accr[i][j], nai, bi[j]);
                    dmma884(acci[i][j], ai, br[j]);
                }
            }
        }
        __syncthreads();  // every warp has left stage kt & 1 before the prefetch of tile kt + 2 overwrites it
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + wm * 32 + i * 8 + g;
        if (m >= p.M) continue;
        const int64_t mo = p.cm.at(m);
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int n = n0 + j * 8 + 2 * t + h;
                if (n >= p.N) continue;
                c128* dst = C + mo + p.cn.at(n);
                const double vr = accr[i][j][h], vi = acci[i][j][h];
                c128 o;
                o.x = p.alpha.x * vr - p.alpha.y * vi;
                o.y = p.alpha.x * vi + p.alpha.y * vr;
                if (!p.beta_zero) {
                    const c128 old = *dst;
                    o.x += p.beta.x * old.x - p.beta.y * old.y;
                    o.y += p.beta.x * old.y + p.beta.y * old.x;
                }
                *dst = o;
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming variant of the thin kernel for a single N tile (N <= 16, K <= 128): persistent CTAs, the gate-sized operand
// parked in shared memory once, and ONE software pipeline over the flattened (M tile, k tile) sequence -- a 3-stage
// cp.async ring that runs on into the next M tile while the current one is computed and stored.  What made the
// per-tile kernel latency-bound is taken off the critical path: the row-offset tables (am for the loads, cm for the
// stores) are fetched one tile ahead into registers, the k- and n-offset tables live in shared memory.
constexpr int ST_STAGES = 3, ST_MAXK = 128;
template <int NT>
struct StreamCfg {
    static constexpr int BNT = 8 * NT, PBT = BNT + 2;
    static size_t smem(int K) {
        const int kp = (K + BK - 1) / BK * BK;
        return (size_t)ST_STAGES * BK * PA * sizeof(c128) + (size_t)kp * PBT * sizeof(c128) + (size_t)(kp + BNT) * sizeof(int64_t);
    }
};

template <int NT>
__global__ void __launch_bounds__(TH_THREADS, NT == 1 ? 4 : 3) gemm_c128_stream_kernel(const GemmArgs p, int mtiles) {
    constexpr int BNT = StreamCfg<NT>::BNT, PBT = StreamCfg<NT>::PBT;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int KT = (p.K + BK - 1) / BK, KP = KT * BK;
    c128* As = reinterpret_cast<c128*>(smem_raw);             // [ST_STAGES][BK][PA]
    c128* Bs = As + (size_t)ST_STAGES * BK * PA;              // [KP][PBT], zero padded
    int64_t* aks = reinterpret_cast<int64_t*>(Bs + (size_t)KP * PBT);  // ak offsets [KP]
    int64_t* cns = aks + KP;                                  // cn offsets [BNT]
    const int tid = threadIdx.x, lane = tid & 31, wm = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int z = blockIdx.z;
    const c128* __restrict__ A = p.A + p.ab.at(z);
    const c128* __restrict__ B = p.B + p.bb.at(z);
    c128* __restrict__ C = p.C + p.cb.at(z);

    for (int k = tid; k < KP; k += TH_THREADS) aks[k] = k < p.K ? p.ak.at(k) : 0;
    if (tid < BNT) cns[tid] = tid < p.N ? p.cn.at(tid) : 0;
    for (int e = tid; e < KP * BNT; e += TH_THREADS) {
        const int k = p.b_kfast ? e % KP : e / BNT, n = p.b_kfast ? e / KP : e % BNT;
        const bool ok = k < p.K && n < p.N;
        cp_async16(Bs + k * PBT + n, ok ? (B + p.bn.at(n) + p.bk.at(k)) : p.B, ok);
    }
    cp_async_commit();  // the oldest group: complete at the first wait below

    constexpr int A_PER = BM * BK / TH_THREADS;  // 8
    int a_ml[A_PER], a_kl[A_PER];
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
        const int e = tid + TH_THREADS * i;
        if (p.a_kfast) {
            a_kl[i] = e % BK;
            a_ml[i] = e / BK;
        } else {
            a_ml[i] = e % BM;
            a_kl[i] = e / BM;
        }
    }
    const int ntl = (mtiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // M tiles of this CTA
    const int total = ntl * KT;
    auto tile_m0 = [&](int j) { return ((int)blockIdx.x + j * (int)gridDim.x) * BM; };
    // row offsets of the loads (8 per thread) and of the stores (4 per thread), one tile ahead
    int64_t am_iss[A_PER], am_nxt[A_PER], cm_cur[4], cm_nxt[4];
    unsigned ok_iss = 0, ok_nxt = 0;
    auto fetch_am = [&](int j, int64_t (&dst)[A_PER], unsigned& okm) {
        okm = 0;
        if (j >= ntl) return;
        const int m0 = tile_m0(j);
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const bool ok = (m0 + a_ml[i]) < p.M;
            dst[i] = ok ? p.am.at(m0 + a_ml[i]) : 0;
            okm |= (ok ? 1u : 0u) << i;
        }
    };
    auto fetch_cm = [&](int j, int64_t (&dst)[4]) {
        if (j >= ntl) return;
        const int m0 = tile_m0(j);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + wm * 32 + i * 8 + g;
            dst[i] = m < p.M ? p.cm.at(m) : -1;
        }
    };
    fetch_am(0, am_iss, ok_iss);
    fetch_am(1, am_nxt, ok_nxt);
    fetch_cm(0, cm_cur);
    fetch_cm(1, cm_nxt);
    __syncthreads();  // aks / cns visible

    int qi = 0, ji = 0, kti = 0;
    auto issue = [&]() {
        if (qi < total) {
            c128* as = As + (size_t)(qi % ST_STAGES) * BK * PA;
#pragma unroll
            for (int i = 0; i < A_PER; ++i) {
                const int kg = kti * BK + a_kl[i];
                const bool ok = ((ok_iss >> i) & 1u) && kg < p.K;
                cp_async16(as + a_kl[i] * PA + a_ml[i], ok ? (A + am_iss[i] + aks[kg]) : p.A, ok);
            }
            if (++kti == KT) {
                kti = 0;
                ++ji;
#pragma unroll
                for (int i = 0; i < A_PER; ++i) am_iss[i] = am_nxt[i];
                ok_iss = ok_nxt;
                fetch_am(ji + 1, am_nxt, ok_nxt);
            }
        }
        cp_async_commit();
        ++qi;
    };
    double accr[4][NT][2], acci[4][NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) accr[i][j][0] = accr[i][j][1] = acci[i][j][0] = acci[i][j][1] = 0.0;
#pragma unroll
    for (int s = 0; s < ST_STAGES - 1; ++s) issue();
    const int sgnA = p.conjA ? 0x80000000 : 0, sgnB = p.conjB ? 0x80000000 : 0;
    int jc = 0, kt = 0;
    for (int qc = 0; qc < total; ++qc) {
        cp_async_wait<ST_STAGES - 2>();
        __syncthreads();  // stage qc landed for everyone; everyone has left stage qc - 1, which the next issue overwrites
        issue();
        const c128* as = As + (size_t)(qc % ST_STAGES) * BK * PA + wm * 32 + g;
        const c128* bs = Bs + (size_t)kt * BK * PBT + g;
#pragma unroll
        for (int kk = 0; kk < BK / 4; ++kk) {
            double br[NT], bi[NT];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const c128 v = bs[(kk * 4 + t) * PBT + j * 8];
                br[j] = v.x;
                bi[j] = flip_sign(v.y, sgnB);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const c128 v = as[(kk * 4 + t) * PA + i * 8];
                const double ar = v.x, ai = flip_sign(v.y, sgnA), nai = flip_sign(ai, 0x80000000);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    dmma884(accr[i][j], ar, br[j]);
                    dmma884(acci[i][j], ar, bi[j]);
                    dmma884(accr[i][j], nai, bi[j]);
                    dmma884(acci[i][j], ai, br[j]);
                }
            }
        }
        if (++kt == KT) {  // the tile is complete: store it, start the next one
            kt = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int64_t mo = cm_cur[i];
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int n = j * 8 + 2 * t + h;
                        const double vr = accr[i][j][h], vi = acci[i][j][h];
                        accr[i][j][h] = acci[i][j][h] = 0.0;
                        if (mo < 0 || n >= p.N) continue;
                        c128* dst = C + mo + cns[n];
                        c128 o;
                        o.x = p.alpha.x * vr - p.alpha.y * vi;
                        o.y = p.alpha.x * vi + p.alpha.y * vr;
                        if (!p.beta_zero) {
                            const c128 old = *dst;
                            o.x += p.beta.x * old.x - p.beta.y * old.y;
                            o.y += p.beta.x * old.y + p.beta.y * old.x;
                        }
                        *dst = o;
                    }
            }
            ++jc;
#pragma unroll
            for (int i = 0; i < 4; ++i) cm_cur[i] = cm_nxt[i];
            fetch_cm(jc + 1, cm_nxt);
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// Dot-product variant: at most 4 x 4 outputs over a long summed dimension (the closing contractions of a circuit
// network: two big tensors sharing almost all their modes).  On the tile kernel every k tile costs a dependent
// offset-table fetch for 1/32 of a tile of useful data; here consecutive threads walk k, keep the <= 16 complex sums
// in registers and a fixed-order block reduction (shuffle tree, then warps in order) leaves one partial per CTA for
// splitk_reduce_kernel -- deterministic, HBM-bound.
constexpr int DOT_THREADS = 256;
__global__ void __launch_bounds__(DOT_THREADS) gemm_c128_dot_kernel(const GemmArgs p) {
    __shared__ double red[DOT_THREADS / 32][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const c128* __restrict__ A = p.A + p.ab.at(0);
    const c128* __restrict__ B = p.B + p.bb.at(0);
    int64_t amo[4], bno[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        amo[i] = i < p.M ? p.am.at(i) : 0;
        bno[i] = i < p.N ? p.bn.at(i) : 0;
    }
    double ar[4][4], ai[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) ar[i][j] = ai[i][j] = 0.0;
    const double sa = p.conjA ? -1.0 : 1.0, sb = p.conjB ? -1.0 : 1.0;
    for (int64_t k = (int64_t)blockIdx.x * DOT_THREADS + tid; k < p.K; k += (int64_t)gridDim.x * DOT_THREADS) {
        const int64_t ako = p.ak.at(k), bko = p.bk.at(k);
        c128 a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = i < p.M ? A[amo[i] + ako] : make_double2(0.0, 0.0);
            b[i] = i < p.N ? B[bko + bno[i]] : make_double2(0.0, 0.0);
            a[i].y *= sa;
            b[i].y *= sb;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ar[i][j] += a[i].x * b[j].x - a[i].y * b[j].y;
                ai[i][j] += a[i].x * b[j].y + a[i].y * b[j].x;
            }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double vr = warp_sum(ar[i][j]), vi = warp_sum(ai[i][j]);
            if (lane == 0) {
                red[warp][2 * (i + 4 * j)] = vr;
                red[warp][2 * (i + 4 * j) + 1] = vi;
            }
        }
    __syncthreads();
    if (tid < 16) {
        const int i = tid & 3, j = tid >> 2;
        if (i < p.M && j < p.N) {
            double vr = 0.0, vi = 0.0;
            for (int w = 0; w < DOT_THREADS / 32; ++w) {
                vr += red[w][2 * tid];
                vi += red[w][2 * tid + 1];
            }
            p.partial[(size_t)blockIdx.x * p.M * p.N + i + (size_t)p.M * j] = make_double2(vr, vi);
        }
    }
}

// sums the split-K partials in a fixed order (deterministic) and applies alpha / beta
__global__ void splitk_reduce_kernel(const GemmArgs p) {
    int64_t total = (int64_t)p.M * p.N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int m = (int)(idx % p.M), n = (int)(idx / p.M);
        double vr = 0.0, vi = 0.0;
        for (int s = 0; s < p.ksplit; ++s) {
            c128 v = p.partial[(size_t)s * total + idx];
            vr += v.x;
            vi += v.y;
        }
        c128* dst = p.C + p.cm.at(m) + p.cn.at(n) + p.cb.at(0);
        c128 o;
        o.x = p.alpha.x * vr - p.alpha.y * vi;
        o.y = p.alpha.x * vi + p.alpha.y * vr;
        if (!p.beta_zero) {
            c128 old = *dst;
            o.x += p.beta.x * old.x - p.beta.y * old.y;
            o.y += p.beta.x * old.y + p.beta.y * old.x;
        }
        *dst = o;
    }
}

__global__ void build_offsets_kernel(const ModeList ml, int64_t total, int64_t* __restrict__ out) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    int64_t rem = idx, off = 0;
    for (int j = 0; j < ml.n; ++j) {
        int64_t c = rem % ml.ext[j];
        rem /= ml.ext[j];
        off += c * ml.stride[j];
    }
    out[idx] = off;
}

int32_t build_offsets(qb200_ctx* ctx, const ModeList& ml, int64_t total, int64_t* out) {
    if (total <= 0) return QB200_OK;
    int threads = 256;
    int64_t blocks = (total + threads - 1) / threads;
    build_offsets_kernel<<<(unsigned)blocks, threads, 0, ctx->stream>>>(ml, total, out);
    QB_LAUNCH_CHECK(ctx);
    return QB200_OK;
}

int32_t init_gemm(qb200_ctx* ctx) {
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_HALF));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEMM_SMEM_HALF));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_thin_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ThinCfg<1>::SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_thin_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ThinCfg<2>::SMEM));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamCfg<1>::smem(ST_MAXK)));
    QB_CUDA(ctx, cudaFuncSetAttribute(gemm_c128_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StreamCfg<2>::smem(ST_MAXK)));
    return QB200_OK;
}

int32_t launch_gemm(qb200_ctx* ctx, const GemmArgs& args_in) {
    if (args_in.M <= 0 || args_in.N <= 0 || args_in.batch <= 0) return QB200_OK;
    GemmArgs args = args_in;
    if (args.N > args.M) {
        // C^T = B^T A^T: put the larger free dimension on the 128-wide side of the tile
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
    if (args.M <= 4 && args.batch == 1 && args.K >= 4096) {  // (N <= M after the orientation) long dot products
        const int nsplit = (int)std::min<int64_t>(2 * ctx->sm_count, (args.K + DOT_THREADS - 1) / DOT_THREADS);
        args.partial = ws.get<c128>((size_t)nsplit * args.M * args.N);
        if (!args.partial) QB_FAIL(ctx, QB200_E_CUDA, "gemm: split-K workspace allocation failed");
        args.ksplit = nsplit;
        gemm_c128_dot_kernel<<<nsplit, DOT_THREADS, 0, ctx->stream>>>(args);
        QB_LAUNCH_CHECK(ctx);
        splitk_reduce_kernel<<<1, 32, 0, ctx->stream>>>(args);
        QB_LAUNCH_CHECK(ctx);
        return QB200_OK;
    }
    {
        // split-K when the output has too few tiles to fill the machine and K is long
        int64_t tiles = (int64_t)((args.M + BM - 1) / BM) * ((args.N + BN - 1) / BN);
        int KT = (args.K + BK - 1) / BK;
        if (args.batch == 1 && tiles * 2 <= ctx->sm_count && KT >= 64) {
            // long dot products (a handful of output tiles): up to 4 CTAs per SM, each k tile costs a dependent
            // offset-table fetch, so short K ranges per CTA matter more than the (tiny) partial buffers
            int want = (int)std::min<int64_t>((tiles * 8 <= ctx->sm_count ? 4 : 1) * (int64_t)ctx->sm_count / tiles, KT / 16);
            if (want > 1) {
                args.partial = ws.get<c128>((size_t)want * args.M * args.N);
                if (!args.partial) QB_FAIL(ctx, QB200_E_CUDA, "gemm: split-K workspace allocation failed");
                args.ksplit = want;
            }
        }
    }
    // gate-sized operand against a long free dimension: the thin kernel (QB200_GEMM_THIN=0: A/B switch)
    static const bool thin_on = [] {
        const char* e = getenv("QB200_GEMM_THIN");
        return !(e && e[0] == '0');
    }();
    // N <= 16: always; N <= 64 while K <= 16 (HBM-bound: the 4M product and the A tile re-read from L2 by the N tiles cost
    // less than the tile kernel's single resident CTA per SM; with longer K that kernel's 3M product wins)
    if (thin_on && args.ksplit == 1 && args.M >= 2048 && (args.N <= 16 || (args.N <= 64 && args.K <= 16))) {
        const int nt = args.N <= 8 ? 1 : 2;
        const int ntiles = (args.N + 8 * nt - 1) / (8 * nt);
        dim3 grid((unsigned)((args.M + BM - 1) / BM) * ntiles, 1, args.batch);
        if (grid.z > 65535) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "gemm grid too large");
        static const bool stream_on = [] {  // QB200_GEMM_STREAM=0: per-tile thin kernel only (A/B switch)
            const char* e = getenv("QB200_GEMM_STREAM");
            return !(e && e[0] == '0');
        }();
        if (stream_on && ntiles == 1 && args.K <= 32) {  // longer K: the per-tile kernel's 4 CTAs per SM win (K = 128: 91 vs 112 us)
            const int mtiles = (args.M + BM - 1) / BM;
            dim3 sgrid((unsigned)std::min(mtiles, (nt == 1 ? 4 : 3) * ctx->sm_count), 1, args.batch);
            if (nt == 1)
                gemm_c128_stream_kernel<1><<<sgrid, TH_THREADS, StreamCfg<1>::smem(args.K), ctx->stream>>>(args, mtiles);
            else
                gemm_c128_stream_kernel<2><<<sgrid, TH_THREADS, StreamCfg<2>::smem(args.K), ctx->stream>>>(args, mtiles);
            QB_LAUNCH_CHECK(ctx);
            return QB200_OK;
        }
        if (nt == 1)
            gemm_c128_thin_kernel<1><<<grid, TH_THREADS, ThinCfg<1>::SMEM, ctx->stream>>>(args, ntiles);
        else
            gemm_c128_thin_kernel<2><<<grid, TH_THREADS, ThinCfg<2>::SMEM, ctx->stream>>>(args, ntiles);
        QB_LAUNCH_CHECK(ctx);
        return QB200_OK;
    }
    // beta = 1 with alpha = +-1 and no split-K: fold C into the accumulators
    if (args.ksplit == 1 && !args.beta_zero && args.beta.x == 1.0 && args.beta.y == 0.0 && args.alpha.y == 0.0 &&
        (args.alpha.x == 1.0 || args.alpha.x == -1.0))
        args.acc_init = 1;
    // short K, many tiles: the 64-row tile with two CTAs per SM (QB200_GEMM_HALF=0: A/B switch)
    static const bool half_on = [] {
        const char* e = getenv("QB200_GEMM_HALF");
        return !(e && e[0] == '0');
    }();
    const bool half = half_on && args.ksplit == 1 && args.K <= 128 && (int64_t)((args.M + 63) / 64) * ((args.N + BN - 1) / BN) * args.batch >= 4 * ctx->sm_count;
    const int tbm = half ? 64 : BM;
    dim3 grid((args.M + tbm - 1) / tbm, (args.N + BN - 1) / BN, args.ksplit > 1 ? args.ksplit : args.batch);
    if (grid.y > 65535 || grid.z > 65535) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "gemm grid too large");
    static const bool gemm_3m = [] {  // QB200_GEMM_3M=0 selects the 4-DMMA complex product
        const char* e = getenv("QB200_GEMM_3M");
        return !(e && e[0] == '0');
    }();
    if (half)
        (gemm_3m ? gemm_c128_kernel<1, 1> : gemm_c128_kernel<0, 1>)<<<grid, 256, GEMM_SMEM_HALF, ctx->stream>>>(args);
    else
        (gemm_3m ? gemm_c128_kernel<1, 0> : gemm_c128_kernel<0, 0>)<<<grid, GEMM_THREADS, GEMM_SMEM, ctx->stream>>>(args);
    QB_LAUNCH_CHECK(ctx);
    if (args.ksplit > 1) {
        int64_t total = (int64_t)args.M * args.N;
        unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 8);
        splitk_reduce_kernel<<<blocks, 256, 0, ctx->stream>>>(args);
        QB_LAUNCH_CHECK(ctx);
    }
    return QB200_OK;
}

}  // namespace qb

// Plain column-major GEMM view used by QR / MPS code.  opX: 0 = N, 1 = T, 2 = C, 3 = conj (no transpose)
int32_t qb_gemm(qb200_ctx* ctx, int opA, int opB, int64_t M, int64_t N, int64_t K, c128 alpha, const c128* A,
                int64_t lda, const c128* B, int64_t ldb, c128 beta, c128* C, int64_t ldc) {
    using namespace qb;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.A = A;
    g.B = B;
    g.C = C;
    bool ta = (opA == 1 || opA == 2), tb = (opB == 1 || opB == 2);
    // A(i,k): N -> i + k*lda ; T -> k + i*lda
    g.am = {nullptr, ta ? lda : 1};
    g.ak = {nullptr, ta ? 1 : lda};
    g.bk = {nullptr, tb ? ldb : 1};
    g.bn = {nullptr, tb ? 1 : ldb};
    g.cm = {nullptr, 1};
    g.cn = {nullptr, ldc};
    g.ab = g.bb = g.cb = {nullptr, 0};
    g.M = (int)M;
    g.N = (int)N;
    g.K = (int)K;
    g.batch = 1;
    g.conjA = (opA >= 2);
    g.conjB = (opB >= 2);
    g.a_kfast = ta;
    g.b_kfast = !tb;
    g.alpha = alpha;
    g.beta = beta;
    g.beta_zero = (beta.x == 0.0 && beta.y == 0.0);
    return launch_gemm(ctx, g);
}
