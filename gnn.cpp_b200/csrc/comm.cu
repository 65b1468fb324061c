// comm.cu — NCCL plumbing for the 1-D row-partitioned multi-GPU path (K10).  One process per GPU; the
// communicator is bootstrapped from a ncclUniqueId the caller distributes (torch.distributed, files, MPI).
// NCCL is resolved at run time (dlopen of libnccl.so.2 — inside a PyTorch process that is the copy torch already
// loaded) so the library also loads on boxes without NCCL; collectives run on the context's stream.
#include <dlfcn.h>

#include "common.cuh"

namespace gnn {

// minimal NCCL ABI (stable since 2.x): opaque comm, 128-byte unique id, enums as ints
struct NcclUniqueId { char internal[128]; };
typedef void *ncclComm_t_;
enum { NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t_ *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t_) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return nullptr;
#define LOAD(field, sym)                                              \
    *(void **)(&api.field) = dlsym(api.handle, sym);                  \
    if (!api.field) { dlclose(api.handle); api.handle = nullptr; return nullptr; }
    LOAD(GetUniqueId, "ncclGetUniqueId");
    LOAD(CommInitRank, "ncclCommInitRank");
    LOAD(CommDestroy, "ncclCommDestroy");
    LOAD(AllGather, "ncclAllGather");
    LOAD(AllReduce, "ncclAllReduce");
    LOAD(GetErrorString, "ncclGetErrorString");
#undef LOAD
    return &api;
}

#define GNN_CHECK_NCCL(api, expr)                                                                 \
    do {                                                                                          \
        int _r = (expr);                                                                          \
        if (_r != 0) {                                                                            \
            gnn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, (api)->GetErrorString(_r)); \
            return 4;                                                                             \
        }                                                                                         \
    } while (0)

} // namespace gnn

using namespace gnn;

extern "C" {

int gnn_comm_unique_id_h(void *id_h) {
    GNN_REQUIRE(id_h, "gnn_comm_unique_id_h: NULL argument");
    NcclApi *api = nccl_api();
    GNN_REQUIRE(api, "gnn_comm_unique_id_h: libnccl.so.2 not found (%s)", dlerror());
    NcclUniqueId id;
    GNN_CHECK_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_h, &id, sizeof(id));
    return 0;
}

int gnn_comm_init(gnn_ctx_t *ctx, const void *id_h, int rank, int world) {
    GNN_REQUIRE(ctx && id_h && world >= 1 && rank >= 0 && rank < world, "gnn_comm_init: bad argument");
    NcclApi *api = nccl_api();
    GNN_REQUIRE(api, "gnn_comm_init: libnccl.so.2 not found (%s)", dlerror());
    GNN_CHECK_CUDA(cudaSetDevice(ctx->device));
    NcclUniqueId id;
    memcpy(&id, id_h, sizeof(id));
    ncclComm_t_ comm = nullptr;
    GNN_CHECK_NCCL(api, api->CommInitRank(&comm, world, id, rank));
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->world = world;
    return 0;
}

int gnn_comm_destroy(gnn_ctx_t *ctx) {
    if (!ctx || !ctx->nccl_comm) return 0;
    NcclApi *api = nccl_api();
    if (api) api->CommDestroy((ncclComm_t_)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return 0;
}

int gnn_allgather_rows(gnn_ctx_t *ctx, const float *send, float *recv, int64_t rows_per_rank, int32_t F) {
    GNN_REQUIRE(ctx && send && recv && rows_per_rank > 0 && F > 0, "gnn_allgather_rows: bad argument");
    const size_t count = (size_t)rows_per_rank * F;
    if (ctx->world == 1) {
        if (send != recv) GNN_CHECK_CUDA(cudaMemcpyAsync(recv, send, count * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    GNN_REQUIRE(ctx->nccl_comm, "gnn_allgather_rows: communicator not initialised (gnn_comm_init)");
    NcclApi *api = nccl_api();
    GNN_CHECK_NCCL(api, api->AllGather(send, recv, count, NCCL_FLOAT32, (ncclComm_t_)ctx->nccl_comm, ctx->stream));
    return 0;
}

int gnn_allreduce_sum(gnn_ctx_t *ctx, float *buf, int64_t n) {
    GNN_REQUIRE(ctx && buf && n > 0, "gnn_allreduce_sum: bad argument");
    if (ctx->world == 1) return 0;
    GNN_REQUIRE(ctx->nccl_comm, "gnn_allreduce_sum: communicator not initialised (gnn_comm_init)");
    NcclApi *api = nccl_api();
    GNN_CHECK_NCCL(api, api->AllReduce(buf, buf, (size_t)n, NCCL_FLOAT32, NCCL_SUM, (ncclComm_t_)ctx->nccl_comm,
                                      ctx->stream));
    return 0;
}

} // extern "C"
