// jlama-net replacement: the TP all-reduce ("combine", JlamaService.java:300-359) over NCCL / NVLink
// instead of gRPC through the coordinator.  NCCL is bound at run time with dlopen so that single-GPU
// use has no NCCL dependency (the same libnccl.so.2 torch already loaded is reused when present).
#include "jl_common.cuh"

#include <dlfcn.h>
#include <string.h>

typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt8 = 0, ncclFloat32 = 7 };
enum { ncclSum = 0 };

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int nccl_load(jl_ctx *ctx) {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return JL_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return jl_set_error(ctx, JL_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
        return jl_set_error(ctx, JL_ERR_NCCL, "libnccl is missing required symbols");
    g_nccl.handle = h;
    return JL_OK;
}

#define JL_NCCL_CHECK(ctx, expr)                                                                                  \
    do {                                                                                                          \
        ncclResult_t _r = (expr);                                                                                 \
        if (_r != 0)                                                                                              \
            return jl_set_error(ctx, JL_ERR_NCCL, "%s failed: %s", #expr,                                        \
                                g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "nccl error");                \
    } while (0)

extern "C" int jl_comm_unique_id(jl_ctx *ctx, uint8_t *id128) {
    if (!ctx || !id128) return JL_ERR_INVALID;
    int rc = nccl_load(ctx);
    if (rc) return rc;
    ncclUniqueId id;
    JL_NCCL_CHECK(ctx, g_nccl.GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return JL_OK;
}

extern "C" int jl_comm_destroy(jl_ctx *ctx);

extern "C" int jl_comm_init(jl_ctx *ctx, const uint8_t *id128, int rank, int world) {
    if (!ctx || !id128 || world < 1 || rank < 0 || rank >= world) return JL_ERR_INVALID;
    int rc = nccl_load(ctx);
    if (rc) return rc;
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (ctx->nccl_comm) jl_comm_destroy(ctx); // a second init replaces the communicator instead of leaking it
    ncclUniqueId id;
    memcpy(id.internal, id128, 128);
    ncclComm_t comm;
    JL_NCCL_CHECK(ctx, g_nccl.CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return JL_OK;
}

// device-buffer all-reduce on `stream` (capturable into a CUDA graph)
int jl_comm_allreduce_dev(jl_ctx *ctx, cudaStream_t stream, float *buf, size_t count) {
    if (ctx->world <= 1) return JL_OK;
    if (!ctx->nccl_comm) return jl_set_error(ctx, JL_ERR_NCCL, "communicator not initialised");
    JL_NCCL_CHECK(ctx, g_nccl.AllReduce(buf, buf, count, ncclFloat32, ncclSum, (ncclComm_t)ctx->nccl_comm, stream));
    return JL_OK;
}

// device-buffer all-gather of raw bytes (the CUDA IPC handles of the in-kernel exchange buffers travel this way)
int jl_comm_allgather_dev(jl_ctx *ctx, cudaStream_t stream, const void *send, void *recv, size_t bytes_per_rank) {
    if (!ctx->nccl_comm || !g_nccl.AllGather) return jl_set_error(ctx, JL_ERR_NCCL, "communicator not initialised");
    JL_NCCL_CHECK(ctx, g_nccl.AllGather(send, recv, bytes_per_rank, ncclInt8, (ncclComm_t)ctx->nccl_comm, stream));
    return JL_OK;
}

extern "C" int jl_comm_allreduce_f32(jl_ctx *ctx, float *host_buf, int64_t count) {
    if (!ctx || !host_buf || count < 0) return JL_ERR_INVALID;
    std::lock_guard<std::mutex> lk(ctx->mu);
    JL_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return JL_OK;
    float *d = (float *)jl_scratch(ctx, 0, (size_t)count * 4);
    if (!d) return JL_ERR_OOM;
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(d, host_buf, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream));
    int rc = jl_comm_allreduce_dev(ctx, ctx->stream, d, (size_t)count);
    if (rc) return rc;
    JL_CUDA_CHECK(ctx, cudaMemcpyAsync(host_buf, d, (size_t)count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    JL_CUDA_CHECK(ctx, cudaStreamSynchronize(ctx->stream));
    return JL_OK;
}

extern "C" int jl_comm_destroy(jl_ctx *ctx) {
    if (!ctx) return JL_ERR_INVALID;
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
        ctx->nccl_comm = nullptr;
    }
    ctx->world = 1;
    ctx->rank = 0;
    return JL_OK;
}
