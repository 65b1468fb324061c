// ctx.cu — context, error state and device storage entry points of the C ABI (include/gnn_c.h).
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace gnn {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
} // namespace gnn

int gnn_ctx::workspace(size_t bytes, void **out) {
    if (bytes > ws_bytes) {
        size_t want = bytes + bytes / 8 + (1 << 20);
        void *p = nullptr;
        GNN_CHECK_CUDA(cudaMallocAsync(&p, want, stream));
        if (ws) GNN_CHECK_CUDA(cudaFreeAsync(ws, stream));
        ws = p;
        ws_bytes = want;
        ws_gen++;
    }
    *out = ws;
    return 0;
}

extern "C" {

int gnn_version(void) { return 100; }
const char *gnn_last_error(void) { return gnn::g_err; }

int gnn_ctx_create(int device, void *stream, gnn_ctx_t **out) {
    GNN_REQUIRE(out != nullptr, "gnn_ctx_create: out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        gnn::set_error("gnn_ctx_create: no CUDA device available (%s); this library has no CPU fallback",
                       e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return 3;
    }
    GNN_REQUIRE(device >= 0 && device < n, "gnn_ctx_create: device %d out of range (%d devices)", device, n);
    GNN_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GNN_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    GNN_REQUIRE(prop.major >= 10, "gnn_ctx_create: device %d is sm_%d%d; kernels are built for sm_100a only", device,
                prop.major, prop.minor);
    {   // keep stream-ordered temporaries (sort buffers, workspace) cached in the pool instead of returning them to
        // the driver at every synchronisation: a second structure build then costs kernel time only
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    gnn_ctx *c = new gnn_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->l2_bytes = (size_t)prop.l2CacheSize;
    if (const char *e = getenv("GNN_SPMM_ALT")) c->spmm_alt = atoi(e);
    if (const char *e = getenv("GNN_SPMM_ASYNC")) c->spmm_async = atoi(e);
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        GNN_CHECK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    *out = c;
    return 0;
}

int gnn_ctx_destroy(gnn_ctx_t *ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->nccl_comm) gnn_comm_destroy(ctx);
    if (ctx->ws) cudaFree(ctx->ws);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return 0;
}

int gnn_ctx_sync(gnn_ctx_t *ctx) {
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
void *gnn_ctx_stream(gnn_ctx_t *ctx) { return (void *)ctx->stream; }
int gnn_ctx_sm_count(gnn_ctx_t *ctx) { return ctx->sm_count; }
int64_t gnn_ctx_launch_count(gnn_ctx_t *ctx) { return ctx->launches; }

// Plain device allocations.  (Serving them from the stream-ordered pool was measured in round 2 and reverted: with
// multi-GB tensors of many different sizes the pool keeps remapping physical memory — the op-node train step of the C++
// surface went from a steady 83 ms to an erratic 120-670 ms.  Callers that allocate per operation cache blocks themselves:
// gnn.cpp_b200/host/device.h keeps freed blocks by size, so a steady-state step allocates nothing.)
int gnn_malloc(gnn_ctx_t *ctx, void **ptr, size_t bytes) {
    GNN_CHECK_CUDA(cudaSetDevice(ctx->device));
    GNN_CHECK_CUDA(cudaMalloc(ptr, bytes ? bytes : 4));
    return 0;
}
int gnn_free(gnn_ctx_t *ctx, void *ptr) {
    if (!ptr) return 0;
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    GNN_CHECK_CUDA(cudaFree(ptr));
    return 0;
}
int gnn_memset(gnn_ctx_t *ctx, void *ptr, int value, size_t bytes) {
    GNN_CHECK_CUDA(cudaMemsetAsync(ptr, value, bytes, ctx->stream));
    return 0;
}
int gnn_memcpy_h2d(gnn_ctx_t *ctx, void *dst, const void *src_h, size_t bytes) {
    GNN_CHECK_CUDA(cudaMemcpyAsync(dst, src_h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
int gnn_memcpy_d2h(gnn_ctx_t *ctx, void *dst_h, const void *src, size_t bytes) {
    GNN_CHECK_CUDA(cudaMemcpyAsync(dst_h, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}
int gnn_memcpy_d2d(gnn_ctx_t *ctx, void *dst, const void *src, size_t bytes) {
    GNN_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

} // extern "C"
