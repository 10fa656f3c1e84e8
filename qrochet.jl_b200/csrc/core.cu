// Context, tensor handles and the plain C-ABI plumbing of libqrochet_b200.
#include <mutex>

#include "common.cuh"
#include "gemm_c128.cuh"

static thread_local std::string g_last_error_noctx;

extern "C" {

int32_t qb200_create(int32_t device, qb200_ctx** out) {
    if (!out) return QB200_E_INVALID;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_last_error_noctx = "no CUDA device available (libqrochet_b200 has no CPU fallback)";
        return QB200_E_CUDA;
    }
    if (device < 0 || device >= ndev) return QB200_E_INVALID;
    qb200_ctx* ctx = new qb200_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return QB200_E_CUDA;
    }
    cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaMallocHost(&ctx->scratch_host, 1 << 16);
    if (ctx->scratch_host) memset(ctx->scratch_host, 0, 1 << 16);  // incl. the asynchronous status word (qb_async_status)
    // keep freed workspace cached in the pool instead of returning it to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    int32_t r = qb::init_gemm(ctx);
    if (r == QB200_OK) r = qb::init_gemm_c64(ctx);
    if (r == QB200_OK) r = qb::init_gemm_c64_tc5(ctx);
    if (r == QB200_OK) r = qb::init_lp_update_tc5(ctx);
    if (r == QB200_OK) r = qb_qr_init(ctx);
    if (r == QB200_OK) r = qb_svd_init(ctx);
    if (r != QB200_OK) {
        g_last_error_noctx = ctx->err;
        qb200_destroy(ctx);
        return r;
    }
    *out = ctx;
    return QB200_OK;
}

int32_t qb200_destroy(qb200_ctx* ctx) {
    if (!ctx) return QB200_E_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (qb200_ctx* w : ctx->workers) {
        cudaStreamSynchronize(w->stream);
        cudaStreamDestroy(w->stream);
        cudaEventDestroy(w->ev0);
        cudaEventDestroy(w->ev1);
        cudaFreeHost(w->scratch_host);
        if (w->ev_block) cudaEventDestroy(w->ev_block);
        for (auto e : w->prof_pool) cudaEventDestroy(e);
        delete w;
    }
    ctx->workers.clear();
    if (ctx->nccl_comm) qb200_comm_destroy(ctx);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->scratch_host) cudaFreeHost(ctx->scratch_host);
    delete ctx;
    return QB200_OK;
}

const char* qb200_last_error(qb200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error_noctx.c_str(); }

int32_t qb200_set_stream(qb200_ctx* ctx, void* s) {
    if (!ctx) return QB200_E_INVALID;
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (s) {
        ctx->stream = (cudaStream_t)s;
        ctx->own_stream = false;
    } else {
        QB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return QB200_OK;
}

int32_t qb200_synchronize(qb200_ctx* ctx) {
    if (!ctx) return QB200_E_INVALID;
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return qb_check_async_status(ctx);
}

int64_t qb200_launch_count(qb200_ctx* ctx) {
    if (!ctx) return -1;
    int64_t n = ctx->launches;
    for (qb200_ctx* w : ctx->workers) n += w->launches;
    return n;
}

int32_t qb200_timer_begin(qb200_ctx* ctx) {
    QB_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    return QB200_OK;
}
int32_t qb200_timer_end(qb200_ctx* ctx, double* ms) {
    QB_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    QB_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
    float f = 0;
    QB_CUDA(ctx, cudaEventElapsedTime(&f, ctx->ev0, ctx->ev1));
    if (ms) *ms = f;
    return QB200_OK;
}

}  // extern "C"

qb200_ctx* qb_worker(qb200_ctx* parent, int index) {
    while ((int)parent->workers.size() <= index) {
        qb200_ctx* w = new qb200_ctx();
        w->device = parent->device;
        w->sm_count = parent->sm_count;
        w->is_worker = true;
        w->prof_on = parent->prof_on;
        cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
        cudaEventCreate(&w->ev0);
        cudaEventCreate(&w->ev1);
        cudaMallocHost(&w->scratch_host, 1 << 16);
        if (w->scratch_host) memset(w->scratch_host, 0, 1 << 16);
        cudaEventCreateWithFlags(&w->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
        parent->workers.push_back(w);
    }
    return parent->workers[index];
}

extern "C" {

// phase profiler: enable/disable, and read back {count, total ms, total work} per phase (resets the records)
int32_t qb200_prof_enable(qb200_ctx* ctx, int32_t on) {
    if (!ctx) return QB200_E_INVALID;
    ctx->prof_on = on != 0;
    for (qb200_ctx* w : ctx->workers) w->prof_on = ctx->prof_on;
    return QB200_OK;
}
int32_t qb200_prof_read(qb200_ctx* ctx, int32_t nphases, int64_t* counts, double* ms, double* work) {
    if (!ctx || !counts || !ms || !work) return QB200_E_INVALID;
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < nphases; ++i) {
        counts[i] = 0;
        ms[i] = 0.0;
        work[i] = 0.0;
    }
    std::vector<qb200_ctx*> all = ctx->workers;
    all.push_back(ctx);
    for (qb200_ctx* c : all) {
        cudaStreamSynchronize(c->stream);
        for (auto& r : c->prof_recs) {
            float f = 0.f;
            cudaEventElapsedTime(&f, r.e0, r.e1);
            if (r.phase < nphases) {
                counts[r.phase]++;
                ms[r.phase] += f;
                work[r.phase] += r.work;
            }
            c->prof_pool.push_back(r.e0);
            c->prof_pool.push_back(r.e1);
        }
        c->prof_recs.clear();
    }
    return QB200_OK;
}

static int32_t make_tensor(qb200_ctx* ctx, int32_t dtype, int32_t rank, const int64_t* ext, void* ptr,
                           qb200_tensor** out) {
    if (!ctx || !out || rank < 0 || rank > QB200_MAX_RANK) QB_FAIL(ctx, QB200_E_INVALID, "bad tensor rank %d", rank);
    if (dtype < QB200_C128 || dtype > QB200_F32) QB_FAIL(ctx, QB200_E_UNSUPPORTED, "unknown dtype %d", dtype);
    if (dtype == QB200_F32 && ptr)
        QB_FAIL(ctx, QB200_E_UNSUPPORTED, "wrapping external Float32 memory is not supported (real vectors are held in FP64)");
    qb200_tensor* t = new qb200_tensor();
    // ComplexF32 tensors are stored natively (float2; contraction on the TF32 tensor path, gemm_c64.cu).  Float32
    // tensors (Schmidt vectors: a few KB) are held widened to FP64, which is what every scaling kernel consumes.
    t->user_dtype = dtype;
    dtype = (dtype == QB200_F32) ? QB200_F64 : dtype;
    t->dtype = dtype;
    t->rank = rank;
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) {
        if (ext[i] < 0) {
            delete t;
            QB_FAIL(ctx, QB200_E_INVALID, "negative extent");
        }
        t->ext[i] = ext[i];
        n *= ext[i];
    }
    t->bytes = (size_t)n * dtype_size(dtype);
    if (ptr) {
        t->data = ptr;
        t->owned = false;
    } else {
        t->owned = true;
        t->data = nullptr;
        cudaError_t e = cudaMallocAsync(&t->data, t->bytes ? t->bytes : 16, ctx->stream);
        if (e != cudaSuccess) {
            delete t;
            QB_FAIL(ctx, QB200_E_CUDA, "cudaMallocAsync(%zu): %s", t->bytes, cudaGetErrorString(e));
        }
    }
    *out = t;
    return QB200_OK;
}

int32_t qb200_tensor_alloc(qb200_ctx* ctx, int32_t dtype, int32_t rank, const int64_t* ext, qb200_tensor** out) {
    return make_tensor(ctx, dtype, rank, ext, nullptr, out);
}
int32_t qb200_tensor_wrap(qb200_ctx* ctx, int32_t dtype, int32_t rank, const int64_t* ext, void* p,
                          qb200_tensor** out) {
    if (!p) QB_FAIL(ctx, QB200_E_INVALID, "null device pointer");
    return make_tensor(ctx, dtype, rank, ext, p, out);
}
int32_t qb200_tensor_free(qb200_ctx* ctx, qb200_tensor* t) {
    if (!t) return QB200_OK;
    if (t->owned && t->data) cudaFreeAsync(t->data, ctx->stream);  // stream-ordered: no device sync
    delete t;
    return QB200_OK;
}
int32_t qb200_tensor_upload(qb200_ctx* ctx, qb200_tensor* t, const void* host) {
    if (!t || !host) QB_FAIL(ctx, QB200_E_INVALID, "null argument");
    if (t->user_dtype != t->dtype) {  // widen float -> double on the host, then one copy
        size_t cnt = (size_t)t->numel() * (t->dtype == QB200_C128 ? 2 : 1);
        std::vector<double> wide(cnt);
        const float* src = (const float*)host;
        for (size_t i = 0; i < cnt; ++i) wide[i] = (double)src[i];
        QB_CUDA(ctx, cudaMemcpyAsync(t->data, wide.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return QB200_OK;
    }
    size_t n = (size_t)t->numel() * dtype_size(t->dtype);
    QB_CUDA(ctx, cudaMemcpyAsync(t->data, host, n, cudaMemcpyHostToDevice, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // host buffer is only borrowed for the call
    return QB200_OK;
}
int32_t qb200_tensor_download(qb200_ctx* ctx, const qb200_tensor* t, void* host) {
    if (!t || !host) QB_FAIL(ctx, QB200_E_INVALID, "null argument");
    if (t->user_dtype != t->dtype) {  // narrow double -> float on the host
        size_t cnt = (size_t)t->numel() * (t->dtype == QB200_C128 ? 2 : 1);
        std::vector<double> wide(cnt);
        QB_CUDA(ctx, cudaMemcpyAsync(wide.data(), t->data, cnt * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        float* dst = (float*)host;
        for (size_t i = 0; i < cnt; ++i) dst[i] = (float)wide[i];
        return QB200_OK;
    }
    size_t n = (size_t)t->numel() * dtype_size(t->dtype);
    QB_CUDA(ctx, cudaMemcpyAsync(host, t->data, n, cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return qb_check_async_status(ctx);
}
int32_t qb200_tensor_rank(const qb200_tensor* t) { return t ? t->rank : -1; }
int32_t qb200_tensor_dtype(const qb200_tensor* t) { return t ? t->user_dtype : -1; }
int64_t qb200_tensor_extent(const qb200_tensor* t, int32_t i) { return (t && i >= 0 && i < t->rank) ? t->ext[i] : -1; }
void* qb200_tensor_data(const qb200_tensor* t) { return t ? t->data : nullptr; }

int32_t qb200_tensor_copy(qb200_ctx* ctx, const qb200_tensor* src, qb200_tensor* dst) {
    if (!src || !dst || src->dtype != dst->dtype || src->numel() != dst->numel())
        QB_FAIL(ctx, QB200_E_INVALID, "copy: shape/dtype mismatch");
    QB_CUDA(ctx, cudaMemcpyAsync(dst->data, src->data, (size_t)src->numel() * dtype_size(src->dtype),
                                 cudaMemcpyDeviceToDevice, ctx->stream));
    return QB200_OK;
}
int32_t qb200_tensor_reshape(qb200_ctx* ctx, qb200_tensor* t, int32_t rank, const int64_t* ext) {
    if (!t || rank < 0 || rank > QB200_MAX_RANK) QB_FAIL(ctx, QB200_E_INVALID, "reshape: bad rank");
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= ext[i];
    if (n != t->numel()) QB_FAIL(ctx, QB200_E_INVALID, "reshape: element count differs");
    t->rank = rank;
    for (int i = 0; i < rank; ++i) t->ext[i] = ext[i];
    return QB200_OK;
}

}  // extern "C"
