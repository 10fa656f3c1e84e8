// NCCL sum of per-rank partial results (the `sum(fetch.(partial_results))` of examples/distributed.jl:101) and the
// one-to-all replication of the inputs (the `@everywhere` broadcast of the leaf tensors, examples/distributed.jl:58-64;
// for batched expectation values: the MPS itself).
// libnccl is resolved at run time (dlopen) so that the library itself has no link-time NCCL dependency;
// when torch has already loaded its bundled NCCL the same copy is reused.
#include <dlfcn.h>

#include "common.cuh"

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclFloat64 = 8, ncclSum = 0 };

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
            api.AllReduce = (decltype(api.AllReduce))dlsym(api.lib, "ncclAllReduce");
            api.Broadcast = (decltype(api.Broadcast))dlsym(api.lib, "ncclBroadcast");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.lib, "ncclGetErrorString");
            api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
            api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
        }
    }
    if (!api.lib || !api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) return nullptr;
    return &api;
}
}  // namespace

// ncclBroadcast of `bytes` device bytes in place (root's content lands on every rank), on the context's stream
int32_t qb_comm_broadcast_bytes(qb200_ctx* ctx, void* dev, size_t bytes, int root) {
    NcclApi* a = nccl();
    if (!a || !a->Broadcast || !ctx || !ctx->nccl_comm) QB_FAIL(ctx, QB200_E_COMM, "communicator not initialised");
    if (bytes == 0) return QB200_OK;
    if (!dev) QB_FAIL(ctx, QB200_E_INVALID, "broadcast: null buffer");
    ncclResult_t r = a->Broadcast(dev, dev, bytes, ncclUint8, root, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) QB_FAIL(ctx, QB200_E_COMM, "ncclBroadcast: %s", a->GetErrorString ? a->GetErrorString(r) : "error");
    return QB200_OK;
}

// ncclGroupStart / ncclGroupEnd around a series of broadcasts: one fused launch instead of one per tensor
int32_t qb_comm_group(qb200_ctx* ctx, bool begin) {
    NcclApi* a = nccl();
    if (!a || !a->GroupStart || !a->GroupEnd) return QB200_OK;  // not available: the calls simply run one by one
    ncclResult_t r = begin ? a->GroupStart() : a->GroupEnd();
    if (r != 0) QB_FAIL(ctx, QB200_E_COMM, "ncclGroup%s: %s", begin ? "Start" : "End", a->GetErrorString ? a->GetErrorString(r) : "error");
    return QB200_OK;
}

extern "C" {

int32_t qb200_comm_unique_id(char id_out[128]) {
    NcclApi* a = nccl();
    if (!a || !id_out) return QB200_E_COMM;
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != 0) return QB200_E_COMM;
    memcpy(id_out, id.internal, 128);
    return QB200_OK;
}

int32_t qb200_comm_init(qb200_ctx* ctx, int32_t nranks, int32_t rank, const char id[128]) {
    NcclApi* a = nccl();
    if (!a) QB_FAIL(ctx, QB200_E_COMM, "libnccl.so.2 could not be loaded");
    if (!ctx || !id || nranks < 1 || rank < 0 || rank >= nranks) QB_FAIL(ctx, QB200_E_INVALID, "comm_init: bad argument");
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    ncclComm_t comm = nullptr;
    QB_CUDA(ctx, cudaSetDevice(ctx->device));
    ncclResult_t r = a->CommInitRank(&comm, nranks, uid, rank);
    if (r != 0) QB_FAIL(ctx, QB200_E_COMM, "ncclCommInitRank: %s", a->GetErrorString ? a->GetErrorString(r) : "error");
    ctx->nccl_comm = comm;
    ctx->comm_rank = rank;
    ctx->comm_nranks = nranks;
    return QB200_OK;
}

int32_t qb200_comm_allreduce_sum(qb200_ctx* ctx, double* host_values, int32_t count) {
    NcclApi* a = nccl();
    if (!a || !ctx || !ctx->nccl_comm) QB_FAIL(ctx, QB200_E_COMM, "communicator not initialised");
    if (count < 0 || count > 4096 || !host_values) QB_FAIL(ctx, QB200_E_INVALID, "allreduce: bad count");
    Workspace ws(ctx);
    double* d = ws.get<double>((size_t)count);
    if (!d) QB_FAIL(ctx, QB200_E_CUDA, "allreduce: workspace allocation failed");
    memcpy(ctx->scratch_host, host_values, sizeof(double) * count);
    QB_CUDA(ctx, cudaMemcpyAsync(d, ctx->scratch_host, sizeof(double) * count, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = a->AllReduce(d, d, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != 0) QB_FAIL(ctx, QB200_E_COMM, "ncclAllReduce: %s", a->GetErrorString ? a->GetErrorString(r) : "error");
    QB_CUDA(ctx, cudaMemcpyAsync(ctx->scratch_host, d, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
    QB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(host_values, ctx->scratch_host, sizeof(double) * count);
    return QB200_OK;
}

int32_t qb200_comm_broadcast(qb200_ctx* ctx, qb200_tensor* t, int32_t root) {
    if (!t) QB_FAIL(ctx, QB200_E_INVALID, "broadcast: null tensor");
    return qb_comm_broadcast_bytes(ctx, t->data, (size_t)t->numel() * dtype_size(t->dtype), root);
}

int32_t qb200_comm_destroy(qb200_ctx* ctx) {
    NcclApi* a = nccl();
    if (a && ctx && ctx->nccl_comm) a->CommDestroy((ncclComm_t)ctx->nccl_comm);
    if (ctx) ctx->nccl_comm = nullptr;
    return QB200_OK;
}

}  // extern "C"
