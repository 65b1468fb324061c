// comm.cu — NCCL plumbing for the 1-D row-partitioned multi-GPU path (K10).  One process per GPU; the
// communicator is bootstrapped from a ncclUniqueId the caller distributes (torch.distributed, files, MPI).
// NCCL is resolved at run time (dlopen of libnccl.so.2 — inside a PyTorch process that is the copy torch already
// loaded) so the library also loads on boxes without NCCL; collectives run on the context's stream.
//
// Peer arena (the aggregation-input exchange): every rank cudaMalloc's one arena, the ranks swap CUDA IPC handles
// and map each other's arenas, and an all-gather becomes P-1 copy-engine pushes of the rank's own block straight
// into the peers' arenas over NVLink/NVSwitch (one stream per peer, no SMs, measured 770 GB/s vs 467 GB/s for
// ncclAllGather between two B200s), each followed by a release-store of a sequence number into the peer's flag
// word; the consumer's compute stream spins (one warp) until every peer's flag reached the sequence number.
// The pushes are asynchronous to the compute stream, so whatever is enqueued between gnn_peer_gather_begin and
// gnn_peer_gather_wait overlaps the transfer.
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

// ---------------------------------------------------------------------------------------------------- peer arena
constexpr int PEER_MAX_WORLD = 16, PEER_MAX_SLOTS = 1024;

__global__ void peer_set_flag_kernel(uint32_t *flag, uint32_t seq) {
    __threadfence_system();
    *reinterpret_cast<volatile uint32_t *>(flag) = seq;
    __threadfence_system();
}
// SM push: every CTA copies a grid-strided slice of the rank's block into the same position of every peer's arena
// (one local read, world-1 remote 16-byte stores over NVLink); the last CTA to finish publishes the sequence number
// in every peer's flag word.  Unlike simultaneous copy-engine pushes in both directions (measured ~105 GB/s per
// direction between two B200s), SM stores keep both directions of the links busy.
struct PeerPushArgs {
    uint4 *dst[PEER_MAX_WORLD];
    uint32_t *flag[PEER_MAX_WORLD];
    int n_dst;
};
__global__ void __launch_bounds__(512)
    peer_push_kernel(const PeerPushArgs a, const uint4 *__restrict__ src, size_t n16, uint32_t seq, unsigned *done) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 v0 = __ldg(src + i), v1 = __ldg(src + i + stride), v2 = __ldg(src + i + 2 * stride),
              v3 = __ldg(src + i + 3 * stride);
        for (int d = 0; d < a.n_dst; d++) {
            __stcs(a.dst[d] + i, v0);
            __stcs(a.dst[d] + i + stride, v1);
            __stcs(a.dst[d] + i + 2 * stride, v2);
            __stcs(a.dst[d] + i + 3 * stride, v3);
        }
    }
    for (; i < n16; i += stride) {
        const uint4 v = __ldg(src + i);
        for (int d = 0; d < a.n_dst; d++) __stcs(a.dst[d] + i, v);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) { // every CTA's stores are fenced before its increment
            *done = 0;
            __threadfence_system();
            for (int d = 0; d < a.n_dst; d++) *reinterpret_cast<volatile uint32_t *>(a.flag[d]) = seq;
            __threadfence_system();
        }
    }
}

// one warp: lane q polls the flags of peer q for `count` consecutive slots until each reached its sequence number
// (bounded: a lost peer traps instead of hanging)
struct PeerWaitArgs {
    uint32_t seq[32];
};
__global__ void peer_wait_flags_kernel(const uint32_t *flags, int world, int rank, int count, const PeerWaitArgs w) {
    const int q = threadIdx.x;
    if (q < world && q != rank) {
        const long long t0 = clock64();
        for (int s = 0; s < count; s++) {
            const volatile uint32_t *f = flags + s * PEER_MAX_WORLD + q;
            while ((int32_t)(*f - w.seq[s]) < 0) {
                __nanosleep(200);
                if (clock64() - t0 > 40000000000LL) __trap(); // ~20 s: a peer that never arrives must not hang the box
            }
        }
    }
    __threadfence_system();
}


// ---- 2-D partition: "rows -> columns" scatter --------------------------------------------------------------------
// The rank's block src[rows, ld] is cut into column slices; destination d receives the slice [c0[d], c0[d] + w[d]) of
// every row as a dense [rows, w[d]] block at dst[d] (its position inside the destination's gathered column-slice
// matrix).  One kernel serves every destination (the own rank included): consecutive threads copy consecutive
// 16-byte vectors of one slice row, so remote stores leave as contiguous w[d]*4-byte runs.  The last CTA to finish
// publishes the sequence number in every remote destination's flag word, exactly like peer_push_kernel.
struct PeerScatterArgs {
    float *dst[PEER_MAX_WORLD];
    uint32_t *flag[PEER_MAX_WORLD]; // nullptr for the own rank
    int32_t c0[PEER_MAX_WORLD], w[PEER_MAX_WORLD];
    // halo-only exchange: destination d needs only the rows list[d][0 .. cnt[d]) of the block (ascending local row ids;
    // they keep their position in the destination's gathered matrix); list[d] == nullptr sends all `rows` rows
    const int32_t *list[PEER_MAX_WORLD];
    int64_t cnt[PEER_MAX_WORLD];
    int n_dst;
};
__global__ void __launch_bounds__(512)
    peer_scatter_kernel(const PeerScatterArgs a, const float *__restrict__ src, int64_t ld, int64_t rows, uint32_t seq,
                        unsigned *done) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int d = 0; d < a.n_dst; d++) {
        const int32_t wv = a.w[d] >> 2;
        if (wv <= 0) continue;
        const float *s0 = src + a.c0[d];
        uint4 *o = reinterpret_cast<uint4 *>(a.dst[d]);
        size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (a.list[d]) { // listed rows only
            const int32_t *__restrict__ lst = a.list[d];
            const size_t n = (size_t)a.cnt[d] * wv;
            for (; i < n; i += stride) {
                const size_t li = i / wv, v = i - li * wv;
                const size_t r = (size_t)__ldg(lst + li);
                __stcs(o + r * wv + v, __ldg(reinterpret_cast<const uint4 *>(s0 + r * ld) + v));
            }
            continue;
        }
        const size_t n = (size_t)rows * wv;
        for (; i + 3 * stride < n; i += 4 * stride) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const size_t e = i + u * stride;
                const size_t r = e / wv;
                v[u] = __ldg(reinterpret_cast<const uint4 *>(s0 + r * ld) + (e - r * wv));
            }
#pragma unroll
            for (int u = 0; u < 4; u++) __stcs(o + i + u * stride, v[u]);
        }
        for (; i < n; i += stride) {
            const size_t r = i / wv;
            __stcs(o + i, __ldg(reinterpret_cast<const uint4 *>(s0 + r * ld) + (i - r * wv)));
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            for (int d = 0; d < a.n_dst; d++)
                if (a.flag[d]) *reinterpret_cast<volatile uint32_t *>(a.flag[d]) = seq;
            __threadfence_system();
        }
    }
}
// publish `seq` in the flag words of a set of peers: everything enqueued before it on the stream (e.g. an aggregation
// kernel whose epilogue stored into the peers' arenas) is complete and fenced first
struct PeerSignalArgs {
    uint32_t *flag[PEER_MAX_WORLD];
    int n;
};
__global__ void peer_signal_kernel(const PeerSignalArgs a, uint32_t seq) {
    __threadfence_system();
    if (threadIdx.x < a.n) *reinterpret_cast<volatile uint32_t *>(a.flag[threadIdx.x]) = seq;
    __threadfence_system();
}
// lane q (bit q of peer_mask set) polls the flag of peer q for one slot
__global__ void peer_wait_mask_kernel(const uint32_t *flags, uint32_t peer_mask, uint32_t seq) {
    const int q = threadIdx.x;
    if (q < PEER_MAX_WORLD && ((peer_mask >> q) & 1u)) {
        const long long t0 = clock64();
        const volatile uint32_t *f = flags + q;
        while ((int32_t)(*f - seq) < 0) {
            __nanosleep(100);
            if (clock64() - t0 > 40000000000LL) __trap();
        }
    }
    __threadfence_system();
}

} // namespace gnn

struct gnn_peer_arena {
    int world = 1, rank = 0;
    size_t bytes = 0;                      // data bytes per rank (flags live behind them)
    char *base[gnn::PEER_MAX_WORLD] = {};  // base[rank] = own allocation, others = IPC mappings
    cudaStream_t push[gnn::PEER_MAX_WORLD] = {}; // copy-engine mode: one stream per peer
    cudaStream_t push_sm = nullptr;             // SM mode: one high-priority stream
    unsigned *done = nullptr;                   // SM mode: CTA completion counter
    // transport of a tile: 1 = SM store kernel (default), 0 = copy engines, 2 = in-place ncclAllGather of the panel
    // on the side stream (the pipelined schedule with NCCL's transport, e.g. NVLS multicast on 8 ranks)
    std::vector<cudaEvent_t> ev_slot;          // transport 2: completion event per slot
    bool sm_ctas_set = false;       // GNN_PEER_CTAS given
    int sm_mode = 1, sm_ctas = 64; // 64 CTAs: 649 GB/s between two B200s (32: 617, 16: 458; ncclAllGather: 467)
    cudaEvent_t ev_ready = nullptr, ev_self = nullptr;
    uint32_t seq[gnn::PEER_MAX_SLOTS] = {};      // last sequence number begun per slot (identical on every rank)
    uint32_t *flags(int r) const { return reinterpret_cast<uint32_t *>(base[r] + bytes); }
};

namespace gnn {

} // namespace gnn


namespace gnn {
// ---- 2-D partition plumbing used by trainer_grid.cu (internal; exercised through gnn_gcn_create_grid) ------------
char *peer_base(gnn_peer_arena *a, int rank) { return a->base[rank]; }

// after the work already enqueued on the context's stream: scatter column slices of src[rows, ld] (see
// peer_scatter_kernel) on the side stream; destination i is rank dst_rank[i], byte offset dst_off[i] of its arena
int peer_scatter_begin(gnn_ctx *ctx, gnn_peer_arena *a, int slot, const float *src, int64_t ld, int64_t rows, int n_dst,
                       const int *dst_rank, const size_t *dst_off, const int32_t *c0, const int32_t *w,
                       const int32_t *const *lists, const int64_t *counts) {
    GNN_REQUIRE(ctx && a && slot >= 0 && slot < PEER_MAX_SLOTS - 1 && n_dst <= PEER_MAX_WORLD, "peer_scatter_begin: bad argument");
    GNN_REQUIRE(((uintptr_t)src & 15) == 0 && (ld & 3) == 0, "peer_scatter_begin: source must be 16-byte aligned with ld %% 4 == 0");
    const uint32_t seq = ++a->seq[slot];
    PeerScatterArgs pa;
    pa.n_dst = n_dst;
    for (int i = 0; i < n_dst; i++) {
        const int r = dst_rank[i];
        GNN_REQUIRE(r >= 0 && r < a->world && (dst_off[i] & 15) == 0 && (w[i] & 3) == 0 && (c0[i] & 3) == 0 &&
                        dst_off[i] + (size_t)rows * w[i] * 4 <= a->bytes,
                    "peer_scatter_begin: destination %d outside the arena or misaligned", i);
        pa.dst[i] = reinterpret_cast<float *>(a->base[r] + dst_off[i]);
        pa.flag[i] = r == a->rank ? nullptr : a->flags(r) + slot * PEER_MAX_WORLD + a->rank;
        pa.c0[i] = c0[i];
        pa.w[i] = w[i];
        pa.list[i] = lists ? lists[i] : nullptr;
        pa.cnt[i] = lists && lists[i] ? counts[i] : rows;
    }
    GNN_CHECK_CUDA(cudaEventRecord(a->ev_ready, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamWaitEvent(a->push_sm, a->ev_ready, 0));
    // Unlike the all-gather pushes of the row partition (64 CTAs: more of them only took SMs from the aggregation running
    // beside them), the scatter mostly runs alone, and the link rate keeps rising with the CTA count (2 B200s, 1 x 2
    // grid, exposed exchange per step: 64 CTAs 2.68 ms, 296: 2.74, 592: 1.93) — 4 CTAs per SM unless GNN_PEER_CTAS says otherwise
    const int ctas = a->sm_ctas_set ? a->sm_ctas : 4 * ctx->sm_count;
    peer_scatter_kernel<<<ctas, 512, 0, a->push_sm>>>(pa, src, ld, rows, seq, a->done);
    GNN_LAUNCHED(ctx);
    // the own rank's slice is a local store of the same kernel: peer_wait_mask(.., after_own_scatter) orders the
    // compute stream after it (not here, so that work enqueued in between overlaps the transfer)
    GNN_CHECK_CUDA(cudaEventRecord(a->ev_self, a->push_sm));
    return 0;
}

// compute stream: publish the next sequence number of `slot` to the peers in peer_mask (bit q = rank q)
int peer_signal(gnn_ctx *ctx, gnn_peer_arena *a, int slot, uint32_t peer_mask) {
    GNN_REQUIRE(ctx && a && slot >= 0 && slot < PEER_MAX_SLOTS - 1, "peer_signal: bad argument");
    const uint32_t seq = ++a->seq[slot];
    PeerSignalArgs sa;
    sa.n = 0;
    for (int q = 0; q < a->world; q++)
        if (((peer_mask >> q) & 1u) && q != a->rank) sa.flag[sa.n++] = a->flags(q) + slot * PEER_MAX_WORLD + a->rank;
    if (sa.n == 0) return 0;
    peer_signal_kernel<<<1, 32, 0, ctx->stream>>>(sa, seq);
    GNN_LAUNCHED(ctx);
    return 0;
}

// compute stream: wait until every peer in peer_mask has published the current sequence number of `slot`
int peer_wait_mask(gnn_ctx *ctx, gnn_peer_arena *a, int slot, uint32_t peer_mask, bool after_own_scatter) {
    GNN_REQUIRE(ctx && a && slot >= 0 && slot < PEER_MAX_SLOTS - 1, "peer_wait_mask: bad argument");
    if (after_own_scatter) GNN_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, a->ev_self, 0)); // latest peer_scatter_begin
    peer_mask &= ~(1u << a->rank);
    if (!peer_mask) return 0;
    peer_wait_mask_kernel<<<1, 32, 0, ctx->stream>>>(a->flags(a->rank) + slot * PEER_MAX_WORLD, peer_mask, a->seq[slot]);
    GNN_LAUNCHED(ctx);
    return 0;
}
} // namespace gnn

using namespace gnn;

int gnn_peer_arena_transport(const gnn_peer_arena_t *a) { return a ? a->sm_mode : -1; }

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


/* ---- peer arena --------------------------------------------------------------------------------------------- */
int gnn_peer_arena_create(gnn_ctx_t *ctx, size_t bytes, gnn_peer_arena_t **out) {
    GNN_REQUIRE(ctx && out && bytes > 0, "gnn_peer_arena_create: bad argument");
    GNN_REQUIRE(ctx->world > 1 && ctx->nccl_comm, "gnn_peer_arena_create: needs an initialised communicator");
    GNN_REQUIRE(ctx->world <= PEER_MAX_WORLD, "gnn_peer_arena_create: world %d > %d", ctx->world, PEER_MAX_WORLD);
    NcclApi *api = nccl_api();
    *out = nullptr;
    const int W = ctx->world, R = ctx->rank;
    bytes = (size_t)round_up((int64_t)bytes, 256);
    const size_t flag_bytes = (size_t)PEER_MAX_SLOTS * PEER_MAX_WORLD * 4;
    gnn_peer_arena *a = new gnn_peer_arena();
    a->world = W; a->rank = R; a->bytes = bytes;
    GNN_CHECK_CUDA(cudaMalloc((void **)&a->base[R], bytes + flag_bytes));
    GNN_CHECK_CUDA(cudaMemsetAsync(a->base[R], 0, bytes + flag_bytes, ctx->stream));
    // swap IPC handles (and a per-rank success word) through the communicator
    struct Msg { cudaIpcMemHandle_t h; int32_t ok; int32_t pad[15]; };
    static_assert(sizeof(Msg) == 128, "ipc message size");
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    mine.ok = cudaIpcGetMemHandle(&mine.h, a->base[R]) == cudaSuccess ? 1 : 0;
    Msg *d_all = nullptr;
    GNN_CHECK_CUDA(cudaMalloc((void **)&d_all, sizeof(Msg) * W));
    GNN_CHECK_CUDA(cudaMemcpyAsync(d_all + R, &mine, sizeof(Msg), cudaMemcpyHostToDevice, ctx->stream));
    GNN_CHECK_NCCL(api, api->AllGather(d_all + R, d_all, sizeof(Msg) / 4, NCCL_FLOAT32, (ncclComm_t_)ctx->nccl_comm,
                                      ctx->stream));
    std::vector<Msg> all(W);
    GNN_CHECK_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(Msg) * W, cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    int ok = 1;
    for (int r = 0; r < W; r++) ok &= all[r].ok;
    for (int r = 0; r < W && ok; r++) {
        if (r == R) continue;
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            break;
        }
        a->base[r] = (char *)p;
    }
    // every rank must agree (a rank that failed to map a peer would otherwise deadlock the others)
    int32_t *d_ok = reinterpret_cast<int32_t *>(d_all);
    float okf = ok ? 0.f : 1.f;
    GNN_CHECK_CUDA(cudaMemcpyAsync(d_ok, &okf, 4, cudaMemcpyHostToDevice, ctx->stream));
    GNN_CHECK_NCCL(api, api->AllReduce(d_ok, d_ok, 1, NCCL_FLOAT32, NCCL_SUM, (ncclComm_t_)ctx->nccl_comm, ctx->stream));
    GNN_CHECK_CUDA(cudaMemcpyAsync(&okf, d_ok, 4, cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_all);
    if (okf != 0.f) { // CUDA IPC / peer access unavailable: the caller falls back to ncclAllGather
        gnn_peer_arena_destroy(ctx, a);
        set_error("gnn_peer_arena_create: CUDA IPC peer mapping unavailable on this box");
        return 5;
    }
    for (int r = 0; r < W; r++)
        if (r != R) GNN_CHECK_CUDA(cudaStreamCreateWithFlags(&a->push[r], cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        GNN_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        GNN_CHECK_CUDA(cudaStreamCreateWithPriority(&a->push_sm, cudaStreamNonBlocking, hi));
        GNN_CHECK_CUDA(cudaMalloc((void **)&a->done, 4));
        GNN_CHECK_CUDA(cudaMemsetAsync(a->done, 0, 4, ctx->stream));
        if (const char *e = getenv("GNN_PEER_COPY")) a->sm_mode = !strcmp(e, "ce") ? 0 : (!strcmp(e, "nccl") ? 2 : 1);
        if (const char *e = getenv("GNN_PEER_CTAS")) { a->sm_ctas = atoi(e) > 0 ? atoi(e) : a->sm_ctas; a->sm_ctas_set = atoi(e) > 0; }
    }
    GNN_CHECK_CUDA(cudaEventCreateWithFlags(&a->ev_ready, cudaEventDisableTiming));
    GNN_CHECK_CUDA(cudaEventCreateWithFlags(&a->ev_self, cudaEventDisableTiming));
    *out = a;
    return 0;
}

int gnn_peer_arena_destroy(gnn_ctx_t *ctx, gnn_peer_arena_t *a) {
    if (!a) return 0;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    for (int r = 0; r < a->world; r++) {
        if (a->push[r]) { cudaStreamSynchronize(a->push[r]); cudaStreamDestroy(a->push[r]); }
        if (r != a->rank && a->base[r]) cudaIpcCloseMemHandle(a->base[r]);
    }
    if (a->push_sm) { cudaStreamSynchronize(a->push_sm); cudaStreamDestroy(a->push_sm); }
    if (a->done) cudaFree(a->done);
    if (a->ev_ready) cudaEventDestroy(a->ev_ready);
    if (a->ev_self) cudaEventDestroy(a->ev_self);
    for (auto e : a->ev_slot)
        if (e) cudaEventDestroy(e);
    // an exporter must not free memory a peer still has mapped: order all ranks (collective) before the free
    NcclApi *api = nccl_api();
    if (ctx && ctx->nccl_comm && api && a->base[a->rank]) {
        float *w = reinterpret_cast<float *>(a->flags(a->rank)) + (PEER_MAX_SLOTS - 1) * PEER_MAX_WORLD;
        api->AllReduce(w, w, 1, NCCL_FLOAT32, NCCL_SUM, (ncclComm_t_)ctx->nccl_comm, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
    }
    if (a->base[a->rank]) cudaFree(a->base[a->rank]);
    delete a;
    return 0;
}

void *gnn_peer_arena_local(gnn_peer_arena_t *a) { return a ? (void *)a->base[a->rank] : nullptr; }

int gnn_peer_gather_begin(gnn_ctx_t *ctx, gnn_peer_arena_t *a, int slot, size_t offset, size_t bytes) {
    GNN_REQUIRE(ctx && a && slot >= 0 && slot < PEER_MAX_SLOTS - 1, "gnn_peer_gather_begin: bad argument");
    GNN_REQUIRE(offset + bytes <= a->bytes && bytes % 16 == 0 && offset % 16 == 0,
                "gnn_peer_gather_begin: range [%zu, +%zu) outside the arena (%zu bytes) or not 16-byte aligned", offset,
                bytes, a->bytes);
    const uint32_t seq = ++a->seq[slot];
    GNN_CHECK_CUDA(cudaEventRecord(a->ev_ready, ctx->stream));
    if (a->sm_mode == 2) { // the tile must be the rank's whole block of a [world][bytes] region (one row block)
        GNN_REQUIRE(offset >= (size_t)a->rank * bytes && offset - (size_t)a->rank * bytes + (size_t)a->world * bytes <= a->bytes,
                    "gnn_peer_gather_begin: nccl transport needs whole-panel tiles");
        NcclApi *api = nccl_api();
        if ((int)a->ev_slot.size() <= slot) a->ev_slot.resize(slot + 1, nullptr);
        if (!a->ev_slot[slot]) GNN_CHECK_CUDA(cudaEventCreateWithFlags(&a->ev_slot[slot], cudaEventDisableTiming));
        GNN_CHECK_CUDA(cudaStreamWaitEvent(a->push_sm, a->ev_ready, 0));
        if (bytes)
            GNN_CHECK_NCCL(api, api->AllGather(a->base[a->rank] + offset, a->base[a->rank] + offset - (size_t)a->rank * bytes,
                                              bytes / 4, NCCL_FLOAT32, (ncclComm_t_)ctx->nccl_comm, a->push_sm));
        GNN_CHECK_CUDA(cudaEventRecord(a->ev_slot[slot], a->push_sm));
        return 0;
    }
    if (a->sm_mode) {
        PeerPushArgs pa;
        pa.n_dst = 0;
        for (int i = 1; i < a->world; i++) {
            const int r = (a->rank + i) % a->world; // staggered so the ranks do not all hit the same peer first
            pa.dst[pa.n_dst] = reinterpret_cast<uint4 *>(a->base[r] + offset);
            pa.flag[pa.n_dst] = a->flags(r) + slot * PEER_MAX_WORLD + a->rank;
            pa.n_dst++;
        }
        GNN_CHECK_CUDA(cudaStreamWaitEvent(a->push_sm, a->ev_ready, 0));
        peer_push_kernel<<<a->sm_ctas, 512, 0, a->push_sm>>>(pa, reinterpret_cast<const uint4 *>(a->base[a->rank] + offset),
                                                            bytes / 16, seq, a->done);
        GNN_LAUNCHED(ctx);
        return 0;
    }
    for (int i = 1; i < a->world; i++) {
        const int r = (a->rank + i) % a->world;
        GNN_CHECK_CUDA(cudaStreamWaitEvent(a->push[r], a->ev_ready, 0));
        GNN_CHECK_CUDA(cudaMemcpyAsync(a->base[r] + offset, a->base[a->rank] + offset, bytes, cudaMemcpyDefault, a->push[r]));
        peer_set_flag_kernel<<<1, 1, 0, a->push[r]>>>(a->flags(r) + slot * PEER_MAX_WORLD + a->rank, seq);
        GNN_LAUNCHED(ctx);
    }
    return 0;
}

int gnn_peer_gather_wait(gnn_ctx_t *ctx, gnn_peer_arena_t *a, int slot, int count) {
    GNN_REQUIRE(ctx && a && slot >= 0 && count >= 1 && slot + count < PEER_MAX_SLOTS, "gnn_peer_gather_wait: bad argument");
    if (a->sm_mode == 2) {
        for (int s = slot; s < slot + count; s++)
            if (s < (int)a->ev_slot.size() && a->ev_slot[s]) GNN_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, a->ev_slot[s], 0));
        return 0;
    }
    for (int s0 = 0; s0 < count; s0 += 32) {
        PeerWaitArgs w;
        const int n = count - s0 < 32 ? count - s0 : 32;
        for (int s = 0; s < n; s++) w.seq[s] = a->seq[slot + s0 + s];
        peer_wait_flags_kernel<<<1, 32, 0, ctx->stream>>>(a->flags(a->rank) + (slot + s0) * PEER_MAX_WORLD, a->world,
                                                         a->rank, n, w);
        GNN_LAUNCHED(ctx);
    }
    return 0;
}

} // extern "C"
