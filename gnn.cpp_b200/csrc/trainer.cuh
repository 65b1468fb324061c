// trainer.cuh — state of the fused GCN trainer shared by trainer.cu (single GPU, 1-D row partition) and
// trainer_grid.cu (2-D rows x feature-columns partition).
#pragma once
#include <algorithm>

#include "common.cuh"


int gnn_peer_arena_transport(const gnn_peer_arena_t *a); // comm.cu: 1 SM store kernel, 0 copy engines, 2 ncclAllGather
namespace gnn {
int colsum(gnn_ctx *ctx, int64_t N, int32_t F, const float *A, int64_t lda, float *out);
// gemm.cu: gnn_gemm_nn plus db = column sums of the result, fused into the tensor-core epilogue when possible (*fused)
int gemm_nn_bias_grad(gnn_ctx *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B,
                      int64_t ldb, float *C, int64_t ldc, const float *mask, int64_t ldm, int precision, float *db,
                      bool *fused);
int softmax_xent_launch(gnn_ctx *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                        int64_t n_total, float *loss, float *dZ, int64_t ldd, float *db, bool may_touch_padding);
int copy2d(gnn_ctx *ctx, float *dst, int64_t ldd, const float *src, int64_t lds, int64_t rows, int32_t cols);
}

struct gnn_gcn {
    const gnn_graph *g = nullptr;
    int32_t L = 0;
    std::vector<int32_t> dims, ld;   // F_l and padded leading dimension (multiple of 4)
    std::vector<char> agg_first;     // per layer (index 1..L)
    int64_t n_loc = 0, n_glob = 0, chunk = 0; // local rows, global nodes, rows per rank (dist)
    bool dist = false;
    // parameter slab: [W_1, b_1, ..., W_L, b_L] ; gradient slab same layout + 1 float (local loss sum / N)
    float *params = nullptr, *grads = nullptr, *vel = nullptr;
    int64_t n_params = 0;
    std::vector<int64_t> w_off, b_off;
    // activations
    std::vector<float *> H;  // H[l], l = 1..L  [n_loc, ld[l]]
    std::vector<float *> M;  // aggregated inputs of AF layers [n_loc, ld[l-1]]
    float *S1 = nullptr, *G0 = nullptr, *G1 = nullptr; // scratch [n_loc, maxld]
    float *AG = nullptr;                                 // all-gather buffer [world*chunk, maxld] (dist, NCCL mode)
    // dist, peer mode: one gather region [world][chunk, ldw] per aggregation of a step inside the peer arena;
    // slot(l, dir) = 2*(l-1) + dir (dir 0 forward, 1 backward); the rank produces its own block in place
    gnn_peer_arena_t *arena = nullptr;
    std::vector<size_t> slot_off;
    int comm_mode = 1; // 1 = peer arena pushes (falls back to 0 when IPC is unavailable), 0 = ncclAllGather
    bool nccl_transport = false; // GNN_PEER_COPY=nccl: tiles travel by in-place ncclAllGather on the side stream
    int32_t panel_cols = 128; // peer mode: column panel width of a gathered matrix (pipelines transfer and SpMM)
    // row blocks of the rank's rows (peer mode; one block otherwise): rows rb_row[i]..rb_row[i+1], with the matching
    // nonzero offsets of the forward (CSR) and backward (CSC) structure
    int n_rb = 1;
    int64_t rb_row[9] = {0}, rb_kf[9] = {0}, rb_kb[9] = {0};
    // 2-D partition (trainer_grid.cu): world = Pr x Pc, this rank = gi * Pc + gj.  The structure `g` then holds the
    // rows of row group gi (grp_rows of them, all columns); activations stay 1-D row-partitioned (n_loc rows).
    bool grid = false;
    int Pr = 1, Pc = 1, gi = 0, gj = 0;
    int64_t grp_rows = 0;
    std::vector<size_t> pc_off, y_off; // per aggregation op: arena offset of the gathered column slice / of the output rows
    std::vector<char> scattered;       // per op: this step's rows -> columns scatter has already been issued
    std::vector<void *> owned;         // cudaMalloc'ed buffers of the grid trainer (the rest lives in the arena)
    // halo-only exchange + interior/boundary overlap (trainer_grid.cu): per destination rank the rows of this rank it
    // needs (device lists, [0] forward structure, [1] backward structure), and the structure's rows split into those
    // that need only the rank's own rows ("interior": aggregated while the exchange is in flight) and the rest
    struct SubCsr { int32_t n = 0; int64_t nnz = 0; int32_t *ptr = nullptr, *idx = nullptr, *rows = nullptr; float *val = nullptr; int32_t max_nnz = 0; };
    bool halo_lists = false, split = false;
    std::vector<int32_t *> send_list[2];
    std::vector<int64_t> send_cnt[2];
    SubCsr sub[2][2];                  // [direction][0 interior, 1 boundary]
    size_t need_off = 0;
    double halo_fraction = 1.0;        // listed rows / (peers x local rows): 1 = every peer needs every row
    std::vector<float *> H_local;      // grid: local H_l buffers of aggregate-first layers (transform-first H_l is an arena region)
    float *Xs[2] = {nullptr, nullptr};                   // double-buffered staged inputs for *_h entry points
    int32_t *ys[2] = {nullptr, nullptr};
    int cur_slot = 0;
    bool prefetched = false;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_uploaded = nullptr, ev_consumed = nullptr;
    const int32_t *last_y = nullptr;
    float *loss_d = nullptr;
    int32_t maxld = 0;
    // options
    int precision = 1, profile = 0; // dense transforms: 1 = 3xTF32 on tcgen05 (falls back per shape), 0 = FP32 FMA
    float momentum = 0.f, dampening = 0.f, weight_decay = 0.f;
    int nesterov = 0;
    int optimizer = 0; // 0 = SGD (nn::SGD), 1 = Adam (nn::Adam)
    // CUDA graph of the whole step for launch-bound (small) problems: -1 auto, 0 off, 1 on.  The step is a fixed
    // launch sequence over preallocated buffers, so it is captured once (after an eager warm-up step that sizes the
    // workspace) and replayed while the arguments stay the same.
    int use_graph = -1;
    cudaGraphExec_t graph_exec = nullptr, graph_exec2 = nullptr; // two entries: the host-buffer path alternates slots
    int64_t graph_launches = 0;
    struct GraphKey {
        const float *X; int64_t ldx; const int32_t *y; float lr; float *loss_d;
        uint64_t ws_gen; // the captured kernels bake ctx->ws pointers in: a reallocated workspace invalidates the graph
        bool operator==(const GraphKey &o) const {
            return X == o.X && ldx == o.ldx && y == o.y && lr == o.lr && loss_d == o.loss_d && ws_gen == o.ws_gen;
        }
    } graph_key = {nullptr, 0, nullptr, 0.f, nullptr, 0}, graph_key2 = {nullptr, 0, nullptr, 0.f, nullptr, 0};
    float beta1 = 0.9f, beta2 = 0.999f, adam_eps = 1e-8f;
    float *adam_m = nullptr, *adam_v = nullptr;
    // ReLU tie-break overrides (gnn_gcn_set_relu_overrides): per hidden layer, entries whose forward value is forced to
    // the positive (tiny) or the zero side before anything consumes H_l — aligns the discontinuous `Z > 0` decision of
    // pre-activations within rounding distance of zero with another implementation's (cross-implementation parity)
    struct Override { const int32_t *rows = nullptr, *cols = nullptr; const uint8_t *positive = nullptr; int64_t n = 0; };
    std::vector<Override> overrides; // index = layer
    const uint8_t *train_mask = nullptr; // device uint8[n_loc]: rows that enter the loss (Data::set_mask TRAIN)
    int64_t n_train = 0;                 // selected rows over the whole graph
    int64_t steps = 0;
    int64_t opt_steps = 0, vel_steps = 0; // optimiser steps taken (Adam bias correction) / steps the momentum buffer has seen
    // stats
    double alg_bytes = 0, gemm_flops = 0;
    int32_t n_spmm = 0;
    // profiling
    struct Span { int cls; cudaEvent_t a, b; int32_t F; double bytes; float ms; };
    std::vector<Span> spans;
    size_t span_used = 0;
    double breakdown[6] = {0, 0, 0, 0, 0, 0};
};

namespace gnn {

enum { CLS_SPMM = 0, CLS_GEMM = 1, CLS_LOSS = 2, CLS_BIAS = 3, CLS_SGD = 4, CLS_OTHER = 5 };

struct Prof {
    gnn_ctx *ctx;
    gnn_gcn *m;
    int idx = -1;
    Prof(gnn_ctx *c, gnn_gcn *mm, int cls, int32_t F = 0, double bytes = 0) : ctx(c), m(mm) {
        if (!m->profile) return;
        if (m->span_used == m->spans.size()) {
            gnn_gcn::Span s;
            s.cls = cls;
            s.F = 0; s.bytes = 0; s.ms = 0;
            cudaEventCreate(&s.a);
            cudaEventCreate(&s.b);
            m->spans.push_back(s);
        }
        idx = (int)m->span_used++;
        m->spans[idx].cls = cls;
        m->spans[idx].F = F;
        m->spans[idx].bytes = bytes;
        cudaEventRecord(m->spans[idx].a, ctx->stream);
    }
    ~Prof() {
        if (idx >= 0) cudaEventRecord(m->spans[idx].b, ctx->stream);
    }
};

// comm.cu: 2-D partition plumbing
char *peer_base(gnn_peer_arena *a, int rank);
int peer_scatter_begin(gnn_ctx *ctx, gnn_peer_arena *a, int slot, const float *src, int64_t ld, int64_t rows, int n_dst,
                       const int *dst_rank, const size_t *dst_off, const int32_t *c0, const int32_t *w,
                       const int32_t *const *lists = nullptr, const int64_t *counts = nullptr);
int peer_signal(gnn_ctx *ctx, gnn_peer_arena *a, int slot, uint32_t peer_mask);
int peer_wait_mask(gnn_ctx *ctx, gnn_peer_arena *a, int slot, uint32_t peer_mask, bool after_own_scatter);
// trainer.cu
int apply_relu_overrides(gnn_ctx *ctx, gnn_gcn *m, int32_t l, int64_t r0, int64_t r1);
// trainer_grid.cu
int forward_grid(gnn_ctx *ctx, gnn_gcn *m, const float *X, int64_t ldx);
int backward_grid(gnn_ctx *ctx, gnn_gcn *m, const float *X, int64_t ldx);
float *dz_buffer_grid(gnn_gcn *m, int32_t l);
void recompute_stats_grid(gnn_gcn *m);

static inline double spmm_alg_bytes(int64_t n_out, int64_t nnz, int32_t F) {
    // B_alg = 4(N+1) + nnz*(8 + 4F) + 4*N*F
    return 4.0 * (n_out + 1) + (double)nnz * (8.0 + 4.0 * F) + 4.0 * n_out * F;
}

} // namespace gnn
