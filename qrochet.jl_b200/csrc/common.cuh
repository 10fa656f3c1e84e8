// Shared internals of libqrochet_b200: context, tensor handles, error plumbing, workspace.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/qrochet_b200.h"

typedef double2 c128;

struct qb200_tensor {
    int32_t dtype;       // storage type on the device: C128, C64 (native float2) or F64
    int32_t user_dtype;  // what the caller asked for: F32 vectors are widened to F64 at upload, narrowed at download
    int32_t rank;
    int64_t ext[QB200_MAX_RANK];
    void* data;
    size_t bytes;  // capacity
    bool owned;
    int64_t numel() const {
        int64_t n = 1;
        for (int i = 0; i < rank; ++i) n *= ext[i];
        return n;
    }
};

struct qb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    int64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 148;
    int last_svd_sweeps = 0;
    int qr_last_dependent = 0;  // numerically dependent columns the last qb_qr_matrix on this context had to complete
    int64_t svd_calls = 0, svd_sweeps = 0;  // totals since creation (this context only; workers are summed by the getter)
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_nranks = 1;  // set by qb200_comm_init
    void* nccl_lib = nullptr;
    double* scratch_host = nullptr;  // pinned, 64 KiB
    // optional phase profiler (CUDA events on the context stream; resolved lazily)
    bool prof_on = false;
    struct ProfRec {
        int phase;
        cudaEvent_t e0, e1;
        double work;  // algorithmic flops (or bytes) of the timed region
    };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    // worker contexts (own stream / pinned scratch) used to run independent bond updates of a TEBD layer
    // concurrently; owned by the parent context
    std::vector<qb200_ctx*> workers;
    bool is_worker = false;
    cudaEvent_t ev_block = nullptr;  // cudaEventBlockingSync: worker threads sleep instead of spinning
};

// host waits for the context's stream.  Worker contexts (one host thread each, up to 12 per process and one process
// per GPU) wait on a blocking event so that idle threads do not burn the host cores the other ranks need.
static inline cudaError_t qb_stream_sync(qb200_ctx* ctx) {
    if (!ctx->is_worker || !ctx->ev_block) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e = cudaEventRecord(ctx->ev_block, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(ctx->ev_block);
}

// Asynchronous failure word: kernels that can time out on a hardware barrier (the tcgen05 GEMM waiting for an MMA
// commit) raise it from the device -- it lives in the context's pinned scratch page, which the device addresses
// directly -- and the next synchronising call (download, synchronize, norm) turns it into an error instead of letting
// wrong numbers through.  The last double of the 64 KiB page is never used for staging.
static inline int* qb_async_status(qb200_ctx* ctx) { return reinterpret_cast<int*>(ctx->scratch_host + 8191); }
static inline int32_t qb_check_async_status(qb200_ctx* ctx) {
    if (!ctx->scratch_host) return 0;
    volatile int* st = qb_async_status(ctx);
    if (*st) {
        *st = 0;
        ctx->err = "a tcgen05 kernel timed out waiting for an MMA commit: results of this stream are invalid";
        return -2;  // QB200_E_CUDA
    }
    return 0;
}

int32_t qb_svd_init(qb200_ctx* ctx);
int32_t qb_qr_init(qb200_ctx* ctx);
qb200_ctx* qb_worker(qb200_ctx* parent, int index);  // creates workers lazily

enum QbPhase {
    QB_PH_THETA_GEMM = 0,
    QB_PH_GATE = 1,
    QB_PH_SVD = 2,
    QB_PH_JGRAM = 3,
    QB_PH_JEVD = 4,
    QB_PH_JUPDATE = 5,
    QB_PH_EMIT = 6,
    QB_PH_SCALE = 7,
    QB_PH_QR = 8,
    QB_PH_TN_GEMM = 9,
    QB_PH_LP_GRAM = 10,    // low-precision (FP32 / TF32) stage of the mixed-precision Jacobi SVD: Gram
    QB_PH_LP_UPDATE = 11,  // ... update of the FP32 shadow [X; V]
    QB_PH_LP_GLUE = 12,    // ... V orthonormalised in FP64 and applied: X <- X0 V
    QB_PH_COUNT = 13
};

// RAII phase timer: records two events on the stream when profiling is enabled, otherwise free
struct PhaseTimer {
    qb200_ctx* ctx;
    int idx = -1;
    PhaseTimer(qb200_ctx* c, int phase, double work) : ctx(c) {
        if (!c->prof_on) return;
        qb200_ctx::ProfRec r;
        r.phase = phase;
        r.work = work;
        for (cudaEvent_t* e : {&r.e0, &r.e1}) {
            if (!c->prof_pool.empty()) {
                *e = c->prof_pool.back();
                c->prof_pool.pop_back();
            } else {
                cudaEventCreate(e);
            }
        }
        cudaEventRecord(r.e0, c->stream);
        c->prof_recs.push_back(r);
        idx = (int)c->prof_recs.size() - 1;
    }
    ~PhaseTimer() {
        if (idx >= 0) cudaEventRecord(ctx->prof_recs[idx].e1, ctx->stream);
    }
};

static inline size_t dtype_size(int32_t dt) {
    switch (dt) {
        case QB200_C128: return 16;
        case QB200_C64: return 8;
        case QB200_F64: return 8;
        case QB200_F32: return 4;
    }
    return 0;
}

#define QB_FAIL(ctx, code, ...)                          \
    do {                                                 \
        char _buf[512];                                  \
        snprintf(_buf, sizeof(_buf), __VA_ARGS__);       \
        if (ctx) (ctx)->err = _buf;                      \
        return (code);                                   \
    } while (0)

#define QB_CUDA(ctx, call)                                                                         \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            QB_FAIL(ctx, QB200_E_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
    } while (0)

#define QB_TRY(call)                 \
    do {                             \
        int32_t _r = (call);         \
        if (_r != QB200_OK) return _r; \
    } while (0)

// QB200_SYNC_DEBUG=1: every checked launch is announced on stderr and waited for (locates a hanging kernel)
static inline bool qb_sync_debug() {
    static const bool on = getenv("QB200_SYNC_DEBUG") != nullptr;
    return on;
}

#define QB_LAUNCH_CHECK(ctx)                                                  \
    do {                                                                      \
        (ctx)->launches++;                                                    \
        QB_CUDA(ctx, cudaGetLastError());                                     \
        if (qb_sync_debug()) {                                                \
            fprintf(stderr, "[qb200 launch] %s:%d ...", __FILE__, __LINE__);  \
            fflush(stderr);                                                   \
            QB_CUDA(ctx, cudaStreamSynchronize((ctx)->stream));               \
            fprintf(stderr, " done\n");                                       \
        }                                                                     \
    } while (0)

// stream-ordered workspace from the CUDA memory pool (cached by the driver pool; no sync on free)
struct Workspace {
    qb200_ctx* ctx;
    std::vector<void*> ptrs;
    explicit Workspace(qb200_ctx* c) : ctx(c) {}
    ~Workspace() {
        for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
    }
    template <typename T>
    T* get(size_t n) {
        void* p = nullptr;
        if (n == 0) n = 1;
        if (cudaMallocAsync(&p, n * sizeof(T), ctx->stream) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (T*)p;
    }
};

// ---- internal entry points shared between translation units -----------------------------------
// plain column-major complex GEMM view: C(MxN, ldc) = alpha * op(A) * op(B) + beta * C
// opA: 0 = N, 1 = T, 2 = C (conj transpose), 3 = conj only (no transpose)
int32_t qb_gemm(qb200_ctx* ctx, int opA, int opB, int64_t M, int64_t N, int64_t K, c128 alpha, const c128* A,
                int64_t lda, const c128* B, int64_t ldb, c128 beta, c128* C, int64_t ldc);

// thin QR of a column-major m x n matrix (ld = lda); Q: m x k (ldq), R: k x n (ldr), k = min(m,n).
// A is not modified.
// passes = 2: Q orthonormal to machine precision; passes = 1: only R is trustworthy; passes = QB_QR_R_ONLY: the caller
// uses R alone (SVD preconditioner) -- R to machine precision, Q as the first pass left it when the refinement applies
constexpr int QB_QR_R_ONLY = -1;
int32_t qb_qr_matrix(qb200_ctx* ctx, int64_t m, int64_t n, const c128* A, int64_t lda, c128* Q, int64_t ldq,
                     c128* R, int64_t ldr, int passes = 2);

struct SvdOut {
    // sigma sorted descending on host (k = min(m,n)), filled by qb_svd_factor
    std::vector<double> sigma;
};
// One-sided block-Jacobi SVD of a column-major m x n matrix A (not modified).
// After the call the factorisation lives in a workspace; qb_svd_emit writes the first `kept`
// singular triplets.  Two-phase so that the truncation decision is taken on the host in between.
struct SvdState;
int32_t qb_svd_factor(qb200_ctx* ctx, int64_t m, int64_t n, const c128* A, int64_t lda, SvdState** st,
                      std::vector<double>& sigma);
// U: m x kept (ldu) ; S: kept doubles (device, may be null) ; V written either as
//   vmode 0: Vc = conj(V) n x kept (ldv)       [Tenet's third factor]
//   vmode 1: Vh = V^H    kept x n (ldv)        [row-major-in-bond layout for the fused MPS path]
// optional fused inverse scales: rows of U multiplied by uinv[row % uinv_len] and columns of V^H
// (index j) multiplied by vinv[j / vinv_div] (both may be null).
int32_t qb_svd_emit(qb200_ctx* ctx, SvdState* st, int64_t kept, c128* U, int64_t ldu, double* S, c128* V,
                    int64_t ldv, int vmode, const double* uinv, int64_t uinv_len, const double* vinv,
                    int64_t vinv_div, double sigma_scale);
void qb_svd_release(qb200_ctx* ctx, SvdState* st);

// Cholesky-QR step on a 64-column panel with the Jacobi gram / update kernels (svd_jacobi.cu), used by K4
// before2_dev (may be null): squared norms the 64 columns had at the start of the Gram-Schmidt pass; a column whose
// squared norm is now below dep_tol^2 times that raises *fail_dev (numerically dependent column, see qr.cu)
int32_t qb_cholqr_panel_step(qb200_ctx* ctx, c128* P, int64_t ld, int64_t m, c128* R, int64_t ldr, c128* Gpart,
                             c128* Wbuf, int* flags_dev, int* fail_dev, const double* before2_dev, double dep_tol);
size_t qb_cholqr_gpart_elems(qb200_ctx* ctx);
namespace qb {
// timing harness of the Jacobi update kernel and its diagnostic variants (svd_jacobi.cu), used by the diagnostics library
int32_t qb_update_bench(qb200_ctx* ctx, int k, int steps, double* us_out, int nvar);
// mixed-precision Jacobi, stage A (jacobi_lp_tc5.cu): S_p <- S_p W_p on the FP32 shadow, tcgen05 3xTF32
int32_t init_lp_update_tc5(qb200_ctx* ctx);
int32_t launch_lp_update_tc5(qb200_ctx* ctx, float2* S, int64_t lds, int64_t rows, int nb, int step, const c128* Wg,
                             const int* flags, int npairs);
}

// (left | right) matricisation of a tensor: returns a column-major rows x cols matrix (a permuted copy in
// `ws` unless `order` is the identity)
int32_t qb_matricize(qb200_ctx* ctx, const qb200_tensor* A, const int32_t* order, int32_t nleft, Workspace& ws,
                     const c128** mat, int64_t* rows, int64_t* cols);

// NCCL broadcast of raw device bytes from `root` (comm.cu); the communicator must have been set up by qb200_comm_init
int32_t qb_comm_broadcast_bytes(qb200_ctx* ctx, void* dev, size_t bytes, int root);
int32_t qb_comm_group(qb200_ctx* ctx, bool begin);  // ncclGroupStart / ncclGroupEnd

// elementwise helpers (elementwise.cu)
int32_t qb_scale_mode_raw(qb200_ctx* ctx, const c128* in, c128* out, int64_t inner, int64_t d, int64_t outer,
                          const double* vec, int inverse, double atol);
int32_t qb_scale_rows_cols(qb200_ctx* ctx, const c128* in, c128* out, int64_t rows, int64_t cols,
                           const double* rvec, int64_t rlen, const double* cvec, int64_t cdiv);
int32_t qb_copy_matrix(qb200_ctx* ctx, int64_t m, int64_t n, const c128* A, int64_t lda, c128* B, int64_t ldb,
                       int conj_transpose);
int32_t qb_apply_gate2(qb200_ctx* ctx, c128* theta, int64_t chil, int64_t chir, const c128* gate_dev);
int32_t qb_apply_gate1(qb200_ctx* ctx, c128* t, int64_t inner, int64_t p, int64_t outer, const c128* gate_dev);
int32_t qb_sumsq(qb200_ctx* ctx, const double* x, int64_t n_doubles, double* result_host);
// ComplexF32 <-> ComplexF64 conversion of `n` dense elements (factorisations of C64 tensors run in FP64)
int32_t qb_widen_c64(qb200_ctx* ctx, const void* src_float2, c128* dst, int64_t n);
int32_t qb_narrow_c128(qb200_ctx* ctx, const c128* src, void* dst_float2, int64_t n);
