// trainer.cu — fused full-batch GCN train step (forward + loss + backward + SGD) over preallocated device
// buffers: the launch sequence a main.cpp loop over graph::GCNConv / nn::cross_entropy_loss / nn::SGD issues
// (SURVEY.md §3.5), as one fixed, CUDA-graph-capturable schedule.
//
// Layer l (1..L):  Z_l = A_hat (H_{l-1} W_l^T) + b_l ,  H_l = ReLU(Z_l) for l < L, logits = Z_L.
// Because A_hat (H W^T) == (A_hat H) W^T, each layer aggregates at the narrower width:
//   TF (transform first, F_l <= F_{l-1}):  P = H W^T ;            H_l = relu(A_hat P + b)      [GEMM, SpMM+epilogue]
//       backward: dP = A_hat^T dZ ; dW = dP^T H_{l-1} ; dZ_{l-1} = (dP W) . [H_{l-1} > 0]
//   AF (aggregate first, F_{l-1} < F_l):   M = A_hat H_{l-1} ;    H_l = relu(M W^T + b)        [SpMM, GEMM+epilogue]
//       backward: dW = dZ^T M ; dM = dZ W ; dZ_{l-1} = (A_hat^T dM) . [H_{l-1} > 0]
//       (for l = 1 no input gradient is needed, so an AF first layer has NO backward aggregation)
// Row-partitioned multi-GPU (graph with n_rows < n_cols): every aggregation input is gathered into a
// global-row-order buffer first; weight/bias gradients and the loss are all-reduced in one slab.  The gather is a
// copy-engine push of the rank's block into every peer's arena over NVLink (comm.cu), started as soon as the
// block is produced and awaited right before the aggregation, so the bias-gradient column sums and the dW GEMMs
// of the backward run while the next aggregation input is in flight (NN GEMM before TN GEMM for that reason).
#include "trainer.cuh"

namespace gnn {

static void recompute_stats(gnn_gcn *m) {
    // statistics for the roofline line: every SpMM of one train step
    const gnn_graph *g = m->g;
    m->alg_bytes = 0; m->gemm_flops = 0; m->n_spmm = 0;
    for (int32_t l = 1; l <= m->L; l++) {
        const int32_t Fi = m->dims[l - 1], Fo = m->dims[l];
        if (m->agg_first[l]) {
            m->alg_bytes += spmm_alg_bytes(m->n_loc, g->nnz, Fi); m->n_spmm++;
            if (l > 1) { m->alg_bytes += spmm_alg_bytes(m->n_loc, g->nnz_t, Fi); m->n_spmm++; }
        } else {
            m->alg_bytes += spmm_alg_bytes(m->n_loc, g->nnz, Fo); m->n_spmm++;
            m->alg_bytes += spmm_alg_bytes(m->n_loc, g->nnz_t, Fo); m->n_spmm++;
        }
        m->gemm_flops += 2.0 * m->n_loc * Fi * Fo * (l > 1 ? 3 : 2);
    }
}

__global__ void relu_override_kernel(float *__restrict__ H, int64_t ld, const int32_t *__restrict__ rows,
                                     const int32_t *__restrict__ cols, const uint8_t *__restrict__ positive, int64_t n,
                                     int64_t r0, int64_t r1) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t r = rows[i];
    if (r < r0 || r >= r1) return;
    float *p = H + r * ld + cols[i];
    *p = positive[i] ? fmaxf(*p, 1e-30f) : 0.f;
}
// rows [r0, r1) of H_l are complete: apply the layer's tie-break overrides before anything reads them
int apply_relu_overrides(gnn_ctx *ctx, gnn_gcn *m, int32_t l, int64_t r0, int64_t r1) {
    if (l >= m->L || (size_t)l >= m->overrides.size() || m->overrides[l].n <= 0 || r1 <= r0) return 0;
    const gnn_gcn::Override &o = m->overrides[l];
    relu_override_kernel<<<(unsigned)ceil_div(o.n, 256), 256, 0, ctx->stream>>>(m->H[l], m->ld[l], o.rows, o.cols, o.positive, o.n, r0, r1);
    GNN_LAUNCHED(ctx);
    return 0;
}

// aggregation input must be visible in global row order: all-gather under row partitioning (NCCL mode)
static int gather_input(gnn_ctx *ctx, gnn_gcn *m, const float *local, int32_t ldw, const float **global_out) {
    if (!m->dist) {
        *global_out = local;
        return 0;
    }
    Prof p(ctx, m, CLS_OTHER);
    GNN_TRY(gnn_allgather_rows(ctx, local, m->AG, m->chunk, ldw));
    *global_out = m->AG;
    return 0;
}

// ---- tiling of a gathered matrix ---------------------------------------------------------------------------------
// Peer mode cuts every gathered matrix (padded width ldw, the rank's n_loc rows) into column panels of at most
// m->panel_cols columns and row blocks of the rank's rows.  Panel p of a region is stored panel-major,
// [world][chunk, w_p], so a tile (row block, panel) of one rank is one contiguous byte range: it is pushed into the
// peers' arenas (comm.cu) the moment it is produced, with its own flag slot.
//   * column panels pipeline a transfer with ITS OWN aggregation: SpMM acts on columns independently, so panel p
//     is aggregated as soon as every rank's panel p has landed while panel p+1 is still in flight;
//   * row blocks pipeline a transfer with the PREVIOUS aggregation: everything between the output of one
//     aggregation and the input of the next (bias/ReLU epilogue, dense transform, ReLU mask) is row-local, so the
//     rows of block rb are transformed and pushed while the aggregation still works on block rb+1.
// Outside peer mode there is one panel and one row block and every view is row-major: the schedule degenerates to
// the plain sequence (NCCL mode inserts an all-gather where peer mode waits).
constexpr int MAX_PANELS = 4, MAX_RB = 8;
struct Panels {
    int n = 1;
    int32_t c0[MAX_PANELS] = {0}, w[MAX_PANELS] = {0};
};
// the split itself (also exported as gnn_partition_panels_h so the host-side plan and its CPU tests use the same rule)
static Panels split_panels(int32_t ldw, int32_t pc) {
    Panels P;
    if (pc <= 0 || (int64_t)pc * MAX_PANELS < ldw) pc = (int32_t)round_up(ceil_div(ldw, MAX_PANELS), 4);
    P.n = 0;
    for (int32_t c = 0; c < ldw; c += pc) {
        P.c0[P.n] = c;
        P.w[P.n] = ldw - c < pc ? ldw - c : pc;
        P.n++;
    }
    return P;
}
static Panels panels_of(const gnn_gcn *m, int32_t ldw) { return split_panels(ldw, m->arena ? m->panel_cols : ldw); }
// columns [c0, c0+f) of a logical [n_loc, F] matrix: pointer to (row 0, column c0) and leading dimension
struct View {
    float *ptr;
    int64_t ld;
    float *row(int64_t r) const { return ptr + r * ld; }
};
static inline int op_of(int32_t l, int dir) { return 2 * (l - 1) + dir; }
static inline int slot_of(int op, int p, int rb) { return (op * MAX_PANELS + p) * MAX_RB + rb; }
static inline float *panel_region(gnn_ctx *ctx, gnn_gcn *m, int op, const Panels &P, int p) {
    return reinterpret_cast<float *>((char *)gnn_peer_arena_local(m->arena) + m->slot_off[op]) +
           (size_t)ctx->world * m->chunk * P.c0[p];
}
static inline View own_view(gnn_ctx *ctx, gnn_gcn *m, int op, const Panels &P, int p) {
    return {panel_region(ctx, m, op, P, p) + (size_t)ctx->rank * m->chunk * P.w[p], P.w[p]};
}
static inline View rowmajor_view(float *base, int64_t ld, const Panels &P, int p) { return {base + P.c0[p], ld}; }
static inline int32_t panel_f(const Panels &P, int p, int32_t F) { return F - P.c0[p] < P.w[p] ? F - P.c0[p] : P.w[p]; }

// push tile (row block rb, panel p) of the rank's block of gather region `op` to every peer
static int push_tile(gnn_ctx *ctx, gnn_gcn *m, int op, const Panels &P, int p, int rb) {
    const int64_t r0 = m->rb_row[rb], r1 = m->rb_row[rb + 1];
    const size_t off = m->slot_off[op] +
                       ((size_t)ctx->world * m->chunk * P.c0[p] + (size_t)ctx->rank * m->chunk * P.w[p] + (size_t)r0 * P.w[p]) * 4;
    // an empty row block still publishes its sequence number (16 bytes of padding keep the range valid)
    size_t bytes = r1 > r0 ? (size_t)(r1 - r0) * P.w[p] * 4 : 0;
    if (m->nccl_transport) bytes = (size_t)m->chunk * P.w[p] * 4; // a collective: every rank contributes a whole chunk block
    return gnn_peer_gather_begin(ctx, m->arena, slot_of(op, p, rb), off, bytes);
}
static int wait_panel(gnn_ctx *ctx, gnn_gcn *m, int op, int p) {
    Prof pr(ctx, m, CLS_OTHER);
    return gnn_peer_gather_wait(ctx, m->arena, slot_of(op, p, 0), m->n_rb);
}
// where panel p of dZ_l is written: a transform-first layer aggregates dZ_l itself, so in peer mode it is produced
// straight into the rank's block of that aggregation's gather region; otherwise a row-major ping-pong buffer
static inline View dz_view(gnn_ctx *ctx, gnn_gcn *m, int32_t l, const Panels &P, int p) {
    if (m->arena && !m->agg_first[l]) return own_view(ctx, m, op_of(l, 1), P, p);
    return rowmajor_view(((m->L - l) & 1) ? m->G1 : m->G0, m->ld[l], P, p);
}

// rows of block rb of the aggregation input of layer l: the transformed features P_l = H_{l-1} W_l^T (transform
// first) or H_{l-1} itself (aggregate first); in peer mode written into the gather region and pushed
static int produce_fwd(gnn_ctx *ctx, gnn_gcn *m, int32_t l, int rb, const float *Hin, int64_t ld_in) {
    const int64_t r0 = m->rb_row[rb], rows = m->rb_row[rb + 1] - r0;
    const int32_t Fi = m->dims[l - 1], Fo = m->dims[l];
    const int op = op_of(l, 0);
    if (m->agg_first[l]) {
        if (!m->arena) return 0;
        const Panels P = panels_of(m, m->ld[l - 1]);
        for (int p = 0; p < P.n; p++) {
            const View own = own_view(ctx, m, op, P, p);
            GNN_TRY(copy2d(ctx, own.row(r0), own.ld, Hin + r0 * ld_in + P.c0[p], ld_in, rows, panel_f(P, p, Fi)));
            GNN_TRY(push_tile(ctx, m, op, P, p, rb));
        }
        return 0;
    }
    const float *W = m->params + m->w_off[l];
    const Panels P = panels_of(m, m->ld[l]);
    for (int p = 0; p < P.n; p++) {
        const View out = m->arena ? own_view(ctx, m, op, P, p) : rowmajor_view(m->S1, m->ld[l], P, p);
        if (rows > 0) {
            Prof pr(ctx, m, CLS_GEMM);
            GNN_TRY(gnn_gemm_nt(ctx, rows, panel_f(P, p, Fo), Fi, Hin + r0 * ld_in, ld_in, W + (int64_t)P.c0[p] * Fi, Fi,
                                out.row(r0), out.ld, nullptr, 0, m->precision));
        }
        if (m->arena) GNN_TRY(push_tile(ctx, m, op, P, p, rb));
    }
    return 0;
}

static int forward(gnn_ctx *ctx, gnn_gcn *m, const float *X, int64_t ldx) {
    if (m->grid) return forward_grid(ctx, m, X, ldx);
    const gnn_graph *g = m->g;
    const float *Hin = X;
    int64_t ld_in = ldx;
    for (int rb = 0; rb < m->n_rb; rb++) GNN_TRY(produce_fwd(ctx, m, 1, rb, X, ldx));
    for (int32_t l = 1; l <= m->L; l++) {
        const int32_t Fi = m->dims[l - 1], Fo = m->dims[l];
        const float *W = m->params + m->w_off[l], *b = m->params + m->b_off[l];
        const int relu = l < m->L;
        const int op = op_of(l, 0);
        const bool af = m->agg_first[l];
        const int32_t Fagg = af ? Fi : Fo, ld_agg = af ? m->ld[l - 1] : m->ld[l];
        float *Yout = af ? m->M[l] : m->H[l]; // aggregation output, leading dimension ld_agg
        const Panels P = panels_of(m, ld_agg);
        for (int p = 0; p < P.n; p++) {
            const float *src = nullptr;
            int64_t ld_src = ld_agg;
            if (m->arena) {
                GNN_TRY(wait_panel(ctx, m, op, p));
                src = panel_region(ctx, m, op, P, p);
                ld_src = P.w[p];
            } else if (af) {
                // the collective sends whole [chunk, ld] blocks: the caller's X holds only n_loc rows (fewer than
                // chunk on the last rank when N % world != 0), so it is staged through an internal chunk-row buffer
                const float *send = Hin;
                if (m->dist && l == 1) {
                    GNN_TRY(copy2d(ctx, m->S1, ld_in, Hin, ld_in, m->n_loc, (int32_t)ld_in));
                    send = m->S1;
                }
                GNN_TRY(gather_input(ctx, m, send, (int32_t)ld_in, &src));
                ld_src = ld_in;
            } else {
                GNN_TRY(gather_input(ctx, m, m->S1, m->ld[l], &src));
            }
            for (int rb = 0; rb < m->n_rb; rb++) {
                const int64_t r0 = m->rb_row[rb], r1 = m->rb_row[rb + 1];
                if (r1 > r0) {
                    Prof pr(ctx, m, CLS_SPMM, panel_f(P, p, Fagg),
                            spmm_alg_bytes(r1 - r0, m->rb_kf[rb + 1] - m->rb_kf[rb], panel_f(P, p, Fagg)));
                    GNN_TRY(spmm_rows_range(ctx, g, 0, (int32_t)r0, (int32_t)r1, m->rb_kf[rb], m->rb_kf[rb + 1], src, ld_src,
                                            panel_f(P, p, Fagg), Yout + r0 * ld_agg + P.c0[p], ld_agg,
                                            af ? nullptr : b + P.c0[p], af ? 0 : relu, nullptr, 0));
                }
                if (p + 1 < P.n) continue;
                // the row block is complete: finish the layer for these rows and feed the next layer's exchange
                if (af && r1 > r0) {
                    Prof pr(ctx, m, CLS_GEMM);
                    GNN_TRY(gnn_gemm_nt(ctx, r1 - r0, Fo, Fi, m->M[l] + r0 * m->ld[l - 1], m->ld[l - 1], W, Fi,
                                        m->H[l] + r0 * m->ld[l], m->ld[l], b, relu, m->precision));
                }
                GNN_TRY(apply_relu_overrides(ctx, m, l, r0, r1));
                if (l < m->L) GNN_TRY(produce_fwd(ctx, m, l + 1, rb, m->H[l], m->ld[l]));
            }
        }
        Hin = m->H[l];
        ld_in = m->ld[l];
    }
    return 0;
}

// dZ_L has been written to dz_view(L, .) (and, in peer mode with a transform-first last layer, its tiles pushed)
static int backward(gnn_ctx *ctx, gnn_gcn *m, const float *X, int64_t ldx) {
    if (m->grid) return backward_grid(ctx, m, X, ldx);
    const gnn_graph *g = m->g;
    bool db_fused = false; // db_l already came out of the epilogue of the GEMM that produced dZ_l
    for (int32_t l = m->L; l >= 1; l--) {
        const int32_t Fi = m->dims[l - 1], Fo = m->dims[l];
        const float *W = m->params + m->w_off[l];
        float *dW = m->grads + m->w_off[l], *db = m->grads + m->b_off[l];
        const float *Hin = l > 1 ? m->H[l - 1] : X;
        const int64_t ld_in = l > 1 ? m->ld[l - 1] : ldx;
        const int op = op_of(l, 1);
        const bool db_done = db_fused;
        db_fused = false;
        const Panels Po = panels_of(m, m->ld[l]);                 // panels of dZ_l (width F_l)
        const Panels Pi = panels_of(m, m->ld[l > 1 ? l - 1 : l]); // panels of dZ_{l-1} (width F_{l-1})
        // dZ_{l-1} feeds a transform-first layer's aggregation: produced panel-major and pushed tile by tile
        const bool next_pm = m->arena && l > 1 && !m->agg_first[l - 1];
        const bool dz_pm = m->arena && !m->agg_first[l];
        for (int p = 0; l < m->L && !db_done && p < (dz_pm ? Po.n : 1); p++) { // db_L comes out of the loss kernel
            Prof pr(ctx, m, CLS_BIAS);
            const View dz = dz_view(ctx, m, l, Po, p);
            GNN_TRY(colsum(ctx, m->n_loc, dz_pm ? panel_f(Po, p, Fo) : Fo, dz.ptr, dz.ld, db + (dz_pm ? Po.c0[p] : 0)));
        }
        if (m->agg_first[l]) {
            const View dZ = dz_view(ctx, m, l, Po, 0); // row-major (an aggregate-first layer does not gather dZ)
            if (l > 1)
                for (int rb = 0; rb < m->n_rb; rb++) { // dM[rows, panel] = dZ[rows] W[:, panel]
                    const int64_t r0 = m->rb_row[rb], rows = m->rb_row[rb + 1] - r0;
                    for (int p = 0; p < Pi.n; p++) {
                        const View dM = m->arena ? own_view(ctx, m, op, Pi, p) : rowmajor_view(m->S1, m->ld[l - 1], Pi, p);
                        if (rows > 0) {
                            Prof pr(ctx, m, CLS_GEMM);
                            GNN_TRY(gnn_gemm_nn(ctx, rows, m->arena ? panel_f(Pi, p, Fi) : Fi, Fo, dZ.row(r0), dZ.ld,
                                                W + Pi.c0[p], Fi, dM.row(r0), dM.ld, nullptr, 0, m->precision));
                        }
                        if (!m->arena) break; // single row-major launch
                        GNN_TRY(push_tile(ctx, m, op, Pi, p, rb));
                    }
                }
            {
                Prof pr(ctx, m, CLS_GEMM); // overlaps the transfer of dM
                GNN_TRY(gnn_gemm_tn(ctx, m->n_loc, Fo, Fi, dZ.ptr, dZ.ld, m->M[l], m->ld[l - 1], dW, Fi, m->precision));
            }
            if (l > 1)
                for (int p = 0; p < Pi.n; p++) {
                    const float *src = nullptr;
                    int64_t ld_src = m->ld[l - 1];
                    if (m->arena) {
                        GNN_TRY(wait_panel(ctx, m, op, p));
                        src = panel_region(ctx, m, op, Pi, p);
                        ld_src = Pi.w[p];
                    } else {
                        GNN_TRY(gather_input(ctx, m, m->S1, m->ld[l - 1], &src));
                    }
                    const View dn = dz_view(ctx, m, l - 1, Pi, p);
                    for (int rb = 0; rb < m->n_rb; rb++) {
                        const int64_t r0 = m->rb_row[rb], r1 = m->rb_row[rb + 1];
                        if (r1 > r0) {
                            const int32_t fw = m->arena ? panel_f(Pi, p, Fi) : Fi;
                            Prof pr(ctx, m, CLS_SPMM, fw, spmm_alg_bytes(r1 - r0, m->rb_kb[rb + 1] - m->rb_kb[rb], fw));
                            GNN_TRY(spmm_rows_range(ctx, g, 1, (int32_t)r0, (int32_t)r1, m->rb_kb[rb], m->rb_kb[rb + 1], src,
                                                    ld_src, m->arena ? panel_f(Pi, p, Fi) : Fi, dn.row(r0), dn.ld, nullptr, 0,
                                                    Hin + r0 * ld_in + Pi.c0[p], ld_in));
                        }
                        if (next_pm) GNN_TRY(push_tile(ctx, m, op_of(l - 1, 1), Pi, p, rb));
                    }
                    if (!m->arena) break;
                }
        } else {
            for (int p = 0; p < Po.n; p++) { // dP[:, panel] = A_hat^T dZ[:, panel]
                const float *src = nullptr;
                int64_t ld_src = m->ld[l];
                if (m->arena) {
                    GNN_TRY(wait_panel(ctx, m, op, p));
                    src = panel_region(ctx, m, op, Po, p);
                    ld_src = Po.w[p];
                } else {
                    GNN_TRY(gather_input(ctx, m, dz_view(ctx, m, l, Po, 0).ptr, m->ld[l], &src));
                }
                for (int rb = 0; rb < m->n_rb; rb++) {
                    const int64_t r0 = m->rb_row[rb], r1 = m->rb_row[rb + 1];
                    if (r1 > r0) {
                        Prof pr(ctx, m, CLS_SPMM, panel_f(Po, p, Fo),
                                spmm_alg_bytes(r1 - r0, m->rb_kb[rb + 1] - m->rb_kb[rb], panel_f(Po, p, Fo)));
                        GNN_TRY(spmm_rows_range(ctx, g, 1, (int32_t)r0, (int32_t)r1, m->rb_kb[rb], m->rb_kb[rb + 1], src, ld_src,
                                                panel_f(Po, p, Fo), m->S1 + r0 * m->ld[l] + Po.c0[p], m->ld[l], nullptr, 0,
                                                nullptr, 0));
                    }
                    if (p + 1 < Po.n || l == 1) continue;
                    // dP rows of this block are complete: dZ_{l-1} for these rows first (its transfer then runs under
                    // the remaining row blocks and the dW GEMM below)
                    for (int p2 = 0; p2 < (next_pm ? Pi.n : 1); p2++) {
                        const View dn = dz_view(ctx, m, l - 1, Pi, p2);
                        if (r1 > r0 && !m->arena && m->n_rb == 1) {
                            // one launch covers all of dZ_{l-1}: its column sums (db_{l-1}) come out of the same epilogue
                            Prof pr(ctx, m, CLS_GEMM);
                            GNN_TRY(gemm_nn_bias_grad(ctx, r1 - r0, Fi, Fo, m->S1 + r0 * m->ld[l], m->ld[l], W, Fi, dn.row(r0),
                                                      dn.ld, Hin + r0 * ld_in, ld_in, m->precision,
                                                      m->grads + m->b_off[l - 1], &db_fused));
                        } else if (r1 > r0) {
                            Prof pr(ctx, m, CLS_GEMM);
                            GNN_TRY(gnn_gemm_nn(ctx, r1 - r0, next_pm ? panel_f(Pi, p2, Fi) : Fi, Fo, m->S1 + r0 * m->ld[l],
                                                m->ld[l], W + (next_pm ? Pi.c0[p2] : 0), Fi, dn.row(r0), dn.ld,
                                                Hin + r0 * ld_in + (next_pm ? Pi.c0[p2] : 0), ld_in, m->precision));
                        }
                        if (next_pm) GNN_TRY(push_tile(ctx, m, op_of(l - 1, 1), Pi, p2, rb));
                    }
                }
            }
            Prof pr(ctx, m, CLS_GEMM);
            GNN_TRY(gnn_gemm_tn(ctx, m->n_loc, Fo, Fi, m->S1, m->ld[l], Hin, ld_in, dW, Fi, m->precision));
        }
    }
    return 0;
}

} // namespace gnn

using namespace gnn;

static void drop_graph(gnn_gcn_t *m) {
    if (m->graph_exec) cudaGraphExecDestroy(m->graph_exec);
    if (m->graph_exec2) cudaGraphExecDestroy(m->graph_exec2);
    m->graph_exec = m->graph_exec2 = nullptr;
}

extern "C" {

int gnn_gcn_create(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t L, const int32_t *dims, gnn_gcn_t **out) {
    GNN_REQUIRE(ctx && g && dims && out && L >= 1, "gnn_gcn_create: bad argument");
    GNN_REQUIRE(g->val && g->colptr && (g->valT || g->symmetric),
                "gnn_gcn_create: graph needs gnn_graph_build_csc + gnn_graph_normalize (or a slice of such a graph)");
    for (int32_t l = 0; l <= L; l++) GNN_REQUIRE(dims[l] > 0, "dims cannot be empty or zero");
    gnn_gcn *m = new gnn_gcn();
    m->g = g;
    m->L = L;
    m->n_loc = g->n_rows;
    m->n_glob = g->n_cols;
    m->dist = g->n_rows != g->n_cols || ctx->world > 1;
    if (m->dist) {
        m->chunk = ceil_div(m->n_glob, ctx->world);
        GNN_REQUIRE(m->n_loc <= m->chunk, "gnn_gcn_create: local rows %lld exceed the partition chunk %lld",
                    (long long)m->n_loc, (long long)m->chunk);
    }
    m->dims.assign(dims, dims + L + 1);
    m->ld.resize(L + 1);
    m->agg_first.assign(L + 1, 0);
    m->w_off.assign(L + 1, 0);
    m->b_off.assign(L + 1, 0);
    for (int32_t l = 0; l <= L; l++) {
        m->ld[l] = (int32_t)round_up(dims[l], 4);
        if (m->ld[l] > m->maxld) m->maxld = m->ld[l];
    }
    int64_t off = 0;
    for (int32_t l = 1; l <= L; l++) {
        m->agg_first[l] = dims[l - 1] < dims[l];
        m->w_off[l] = off; off += (int64_t)dims[l] * dims[l - 1];
        m->b_off[l] = off; off += dims[l];
        off = round_up(off, 4); // keep every W 16-byte aligned
    }
    m->n_params = off;
    const int64_t rows_alloc = m->dist ? m->chunk : m->n_loc; // gather sends whole chunks
    auto alloc = [&](float **p, int64_t n) -> int {
        GNN_CHECK_CUDA(cudaMalloc((void **)p, (size_t)n * 4));
        GNN_CHECK_CUDA(cudaMemsetAsync(*p, 0, (size_t)n * 4, ctx->stream));
        return 0;
    };
    GNN_TRY(alloc(&m->params, m->n_params));
    GNN_TRY(alloc(&m->grads, m->n_params + 4));
    m->H.assign(L + 1, nullptr);
    m->M.assign(L + 1, nullptr);
    for (int32_t l = 1; l <= L; l++) {
        GNN_TRY(alloc(&m->H[l], rows_alloc * m->ld[l]));
        if (m->agg_first[l]) GNN_TRY(alloc(&m->M[l], rows_alloc * m->ld[l - 1]));
    }
    GNN_TRY(alloc(&m->S1, rows_alloc * m->maxld));
    GNN_TRY(alloc(&m->G0, rows_alloc * m->maxld));
    GNN_TRY(alloc(&m->G1, rows_alloc * m->maxld));
    if (const char *e = getenv("GNN_COMM")) m->comm_mode = strcmp(e, "nccl") ? 1 : 0; // ablation switches
    // (8 ranks: panels of 32 / 64 / 128 columns all give 14.8-14.9 ms per step, so one width serves every world size)
    if (const char *e = getenv("GNN_PANEL_COLS")) m->panel_cols = (int32_t)round_up(atoi(e) > 0 ? atoi(e) : 1 << 20, 4);
    if (m->dist && ctx->world > 1 && m->comm_mode == 1 && 2 * L * MAX_PANELS * MAX_RB < 1000) {
        // one region per aggregation of a step, wide enough for either layer order
        size_t off = 0;
        m->slot_off.assign(2 * L, 0);
        for (int32_t l = 1; l <= L; l++)
            for (int dir = 0; dir < 2; dir++) {
                const int32_t w = m->ld[l - 1] > m->ld[l] ? m->ld[l - 1] : m->ld[l];
                m->slot_off[2 * (l - 1) + dir] = off;
                off += (size_t)round_up((int64_t)ctx->world * m->chunk * w * 4, 256);
            }
        if (gnn_peer_arena_create(ctx, off, &m->arena) != 0) m->arena = nullptr; // collective; falls back to NCCL
    }
    {
        // Measured on 2 and 4 B200s (products-shaped): splitting the aggregation into row blocks costs more SpMM
        // efficiency (smaller launches, more tails) than the earlier start of the next exchange wins back
        // (4 GPUs: 19.1 ms with 1 block, 19.8 with 2, 20.5 with 4), so one block is the default.
        int want = 1;
        if (const char *e = getenv("GNN_ROW_BLOCKS")) want = atoi(e);
        if (const char *e = getenv("GNN_PEER_COPY"))
            if (!strcmp(e, "nccl")) { want = 1; m->nccl_transport = true; } // an in-place all-gather moves whole panels
        m->n_rb = m->arena ? (want < 1 ? 1 : (want > MAX_RB ? MAX_RB : want)) : 1;
        const int64_t per = round_up(ceil_div(m->n_loc > 0 ? m->n_loc : 1, m->n_rb), 128);
        const bool alias = g->symmetric;
        const int32_t *tptr = alias ? g->rowptr : g->colptr;
        for (int i = 0; i <= m->n_rb; i++) {
            m->rb_row[i] = per * i < m->n_loc ? per * i : m->n_loc;
            int32_t kf = 0, kb = 0;
            GNN_CHECK_CUDA(cudaMemcpyAsync(&kf, g->rowptr + m->rb_row[i], 4, cudaMemcpyDeviceToHost, ctx->stream));
            GNN_CHECK_CUDA(cudaMemcpyAsync(&kb, tptr + m->rb_row[i], 4, cudaMemcpyDeviceToHost, ctx->stream));
            GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
            m->rb_kf[i] = kf;
            m->rb_kb[i] = kb;
        }
    }
    if (m->dist && !m->arena) GNN_TRY(alloc(&m->AG, (int64_t)ctx->world * m->chunk * m->maxld));
    GNN_TRY(alloc(&m->loss_d, 4));
    recompute_stats(m);
    *out = m;
    return 0;
}

int gnn_gcn_destroy(gnn_ctx_t *ctx, gnn_gcn_t *m) {
    if (!m) return 0;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    drop_graph(m);
    cudaFree(m->vel); cudaFree(m->adam_m); cudaFree(m->adam_v);
    if (m->grid) { // activations live in the arena; everything cudaMalloc'ed is listed in `owned`
        for (auto p : m->owned) cudaFree(p);
        for (int dir = 0; dir < 2; dir++) {
            for (auto p : m->send_list[dir]) cudaFree(p);
            for (int k = 0; k < 2; k++) {
                cudaFree(m->sub[dir][k].ptr); cudaFree(m->sub[dir][k].idx); cudaFree(m->sub[dir][k].rows); cudaFree(m->sub[dir][k].val);
            }
        }
        m->loss_d = nullptr;
    } else {
        cudaFree(m->params); cudaFree(m->grads);
        for (auto p : m->H) cudaFree(p);
        for (auto p : m->M) cudaFree(p);
        cudaFree(m->S1); cudaFree(m->G0); cudaFree(m->G1); cudaFree(m->AG);
    }
    if (m->arena) gnn_peer_arena_destroy(ctx, m->arena);
    for (int i = 0; i < 2; i++) { cudaFree(m->Xs[i]); cudaFree(m->ys[i]); }
    if (m->copy_stream) { cudaStreamDestroy(m->copy_stream); cudaEventDestroy(m->ev_uploaded); cudaEventDestroy(m->ev_consumed); }
    cudaFree(m->loss_d);
    for (auto &s : m->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    delete m;
    return 0;
}

int gnn_gcn_set_params_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, const float *W_h, const float *b_h) {
    GNN_REQUIRE(ctx && m && layer >= 1 && layer <= m->L, "gnn_gcn_set_params_h: bad layer");
    const int64_t nw = (int64_t)m->dims[layer] * m->dims[layer - 1];
    if (W_h) GNN_CHECK_CUDA(cudaMemcpyAsync(m->params + m->w_off[layer], W_h, (size_t)nw * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (b_h) GNN_CHECK_CUDA(cudaMemcpyAsync(m->params + m->b_off[layer], b_h, (size_t)m->dims[layer] * 4, cudaMemcpyHostToDevice, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int copy_out(gnn_ctx_t *ctx, float *dst_h, const float *src, size_t n) {
    if (!dst_h) return 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(dst_h, src, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return 0;
}

int gnn_gcn_get_params_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, float *W_h, float *b_h) {
    GNN_REQUIRE(ctx && m && layer >= 1 && layer <= m->L, "gnn_gcn_get_params_h: bad layer");
    GNN_TRY(copy_out(ctx, W_h, m->params + m->w_off[layer], (size_t)m->dims[layer] * m->dims[layer - 1]));
    GNN_TRY(copy_out(ctx, b_h, m->params + m->b_off[layer], (size_t)m->dims[layer]));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int gnn_gcn_get_grads_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, float *dW_h, float *db_h) {
    GNN_REQUIRE(ctx && m && layer >= 1 && layer <= m->L, "gnn_gcn_get_grads_h: bad layer");
    GNN_TRY(copy_out(ctx, dW_h, m->grads + m->w_off[layer], (size_t)m->dims[layer] * m->dims[layer - 1]));
    GNN_TRY(copy_out(ctx, db_h, m->grads + m->b_off[layer], (size_t)m->dims[layer]));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int gnn_gcn_get_activation_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, float *out_h) {
    GNN_REQUIRE(ctx && m && out_h && layer >= 1 && layer <= m->L, "gnn_gcn_get_activation_h: bad layer");
    GNN_CHECK_CUDA(cudaMemcpy2DAsync(out_h, (size_t)m->dims[layer] * 4, m->H[layer], (size_t)m->ld[layer] * 4,
                                     (size_t)m->dims[layer] * 4, (size_t)m->n_loc, cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int gnn_gcn_get_dlogits_h(gnn_ctx_t *ctx, gnn_gcn_t *m, float *out_h) {
    GNN_REQUIRE(ctx && m && out_h, "gnn_gcn_get_dlogits_h: NULL argument");
    GNN_REQUIRE(m->last_y, "gnn_gcn_get_dlogits_h: run gnn_gcn_train_step first");
    // the dZ_L buffer is recycled by the backward ping-pong: recompute it from the kept logits
    const int32_t C = m->dims[m->L];
    const size_t n = (size_t)m->n_loc * m->ld[m->L];
    float *tmp = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&tmp, (n + 4) * 4, ctx->stream));
    GNN_TRY(gnn_softmax_xent(ctx, m->n_loc, C, m->H[m->L], m->ld[m->L], m->last_y, m->n_glob, tmp + n, tmp, m->ld[m->L]));
    GNN_CHECK_CUDA(cudaMemcpy2DAsync(out_h, (size_t)C * 4, tmp, (size_t)m->ld[m->L] * 4, (size_t)C * 4, (size_t)m->n_loc,
                                     cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    GNN_CHECK_CUDA(cudaFreeAsync(tmp, ctx->stream));
    return 0;
}

int gnn_gcn_set_option(gnn_gcn_t *m, const char *key, double value) {
    GNN_REQUIRE(m && key, "gnn_gcn_set_option: NULL argument");
    drop_graph(m); // a captured step bakes the options in
    if (!strcmp(key, "precision")) m->precision = (int)value;
    else if (!strcmp(key, "profile")) m->profile = (int)value;
    else if (!strcmp(key, "momentum")) m->momentum = (float)value;
    else if (!strcmp(key, "dampening")) m->dampening = (float)value;
    else if (!strcmp(key, "weight_decay")) m->weight_decay = (float)value;
    else if (!strcmp(key, "nesterov")) m->nesterov = (int)value;
    else if (!strcmp(key, "optimizer")) m->optimizer = (int)value;
    else if (!strcmp(key, "cuda_graph")) m->use_graph = (int)value;
    else if (!strcmp(key, "beta1")) m->beta1 = (float)value;
    else if (!strcmp(key, "beta2")) m->beta2 = (float)value;
    else if (!strcmp(key, "eps")) m->adam_eps = (float)value;

    else if (!strcmp(key, "agg_first_mask")) { // bit l-1 set -> layer l aggregates first (tests / ablation)
        for (int32_t l = 1; l <= m->L; l++) m->agg_first[l] = (((int64_t)value) >> (l - 1)) & 1;
        if (m->grid) {
            for (int32_t l = 1; l <= m->L; l++) m->H[l] = m->agg_first[l] ? m->H_local[l] : m->M[l]; // M[l] = forward output region
            recompute_stats_grid(m);
        } else {
            recompute_stats(m);
        }
    } else {
        set_error("gnn_gcn_set_option: unknown key '%s'", key);
        return 2;
    }
    return 0;
}

static int ensure_buffers(gnn_ctx_t *ctx, gnn_gcn_t *m) {
    const int64_t rows_alloc = m->dist ? m->chunk : m->n_loc;
    auto zeroed = [&](float **p) -> int {
        if (*p) return 0;
        GNN_CHECK_CUDA(cudaMalloc((void **)p, (size_t)m->n_params * 4));
        GNN_CHECK_CUDA(cudaMemsetAsync(*p, 0, (size_t)m->n_params * 4, ctx->stream));
        return 0;
    };
    if (m->optimizer == 0 && m->momentum != 0.f && !m->vel) { GNN_TRY(zeroed(&m->vel)); m->vel_steps = 0; }
    if (m->optimizer == 1) { GNN_TRY(zeroed(&m->adam_m)); GNN_TRY(zeroed(&m->adam_v)); }
    for (int32_t l = 1; l <= m->L; l++)
        if (m->agg_first[l] && !m->M[l]) {
            GNN_CHECK_CUDA(cudaMalloc((void **)&m->M[l], (size_t)rows_alloc * m->ld[l - 1] * 4));
            GNN_CHECK_CUDA(cudaMemsetAsync(m->M[l], 0, (size_t)rows_alloc * m->ld[l - 1] * 4, ctx->stream));
        }
    return 0;
}

int gnn_gcn_forward(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X, int64_t ldx) {
    GNN_REQUIRE(ctx && m && X && ldx >= m->dims[0], "gnn_gcn_forward: bad argument");
    GNN_TRY(ensure_buffers(ctx, m));
    GNN_REQUIRE(!m->dist || ldx == m->ld[0], "gnn_gcn_forward: row-partitioned mode needs ldx == round_up(F0,4)");
    GNN_TRY(forward(ctx, m, X, ldx));
    // peer mode: a gather region may only be rewritten once every rank is done reading it; the train step gets
    // that ordering from its gradient all-reduce, a forward-only call from this one-word all-reduce
    if (m->arena) GNN_TRY(gnn_allreduce_sum(ctx, m->grads + m->n_params + 1, 1));
    return 0;
}

// the launch sequence of one step (forward, loss, backward, gradient exchange, optimiser, loss copy)
static int train_step_body(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X, int64_t ldx, const int32_t *y, float lr,
                           float *loss_d) {
    GNN_TRY(forward(ctx, m, X, ldx));
    const int32_t C = m->dims[m->L];
    float *loss_slot = m->grads + m->n_params; // rides along with the gradient all-reduce
    {
        Prof p(ctx, m, CLS_LOSS);
        // dZ_L goes where the backward expects it; a panel-major destination with more than one panel is staged
        // through the row-major buffer
        const Panels PL = panels_of(m, m->ld[m->L]);
        const bool pm = m->arena && !m->agg_first[m->L] && !m->grid;
        const View dz = m->grid ? View{dz_buffer_grid(m, m->L), m->ld[m->L]}
                                : ((pm && PL.n > 1) ? View{m->G0, m->ld[m->L]} : dz_view(ctx, m, m->L, PL, 0));
        if (m->train_mask) { // loss over the training nodes only
            GNN_TRY(gnn_softmax_xent_masked(ctx, m->n_loc, C, m->H[m->L], m->ld[m->L], y, m->train_mask, m->n_train,
                                            loss_slot, dz.ptr, dz.ld));
            GNN_TRY(colsum(ctx, m->n_loc, C, dz.ptr, dz.ld, m->grads + m->b_off[m->L]));
        } else {
            // db_L (column sums of dZ_L) is produced by the same kernel from the tiles it already holds
            GNN_TRY(softmax_xent_launch(ctx, m->n_loc, C, m->H[m->L], m->ld[m->L], y, m->n_glob, loss_slot, dz.ptr, dz.ld,
                                        m->grads + m->b_off[m->L], true));
        }
        if (pm)
            for (int p = 0; p < PL.n; p++) {
                if (PL.n > 1) {
                    const View own = own_view(ctx, m, op_of(m->L, 1), PL, p);
                    GNN_TRY(copy2d(ctx, own.ptr, own.ld, m->G0 + PL.c0[p], m->ld[m->L], m->n_loc, panel_f(PL, p, C)));
                }
                for (int rb = 0; rb < m->n_rb; rb++) GNN_TRY(push_tile(ctx, m, op_of(m->L, 1), PL, p, rb));
            }
    }
    GNN_TRY(backward(ctx, m, X, ldx));
    if (m->dist) {
        Prof p(ctx, m, CLS_OTHER);
        GNN_TRY(gnn_allreduce_sum(ctx, m->grads, m->n_params + 1));
    }
    // optimiser state is allocated by ensure_buffers (never inside a capture)
    if (lr != 0.f && m->optimizer == 1) {
        Prof p(ctx, m, CLS_SGD);
        GNN_TRY(gnn_adam_step(ctx, m->n_params, m->params, m->grads, m->adam_m, m->adam_v, lr, m->beta1, m->beta2,
                              m->adam_eps, m->weight_decay, m->opt_steps + 1));
    } else if (lr != 0.f) {
        Prof p(ctx, m, CLS_SGD);
        // torch.optim.SGD: the momentum buffer is INITIALISED with the first gradient it sees (no dampening)
        GNN_TRY(gnn_sgd_step(ctx, m->n_params, m->params, m->grads, m->momentum != 0.f ? m->vel : nullptr, lr, m->momentum,
                             m->dampening, m->weight_decay, m->nesterov, m->vel_steps == 0));
    }
    if (loss_d) GNN_CHECK_CUDA(cudaMemcpyAsync(loss_d, loss_slot, 4, cudaMemcpyDeviceToDevice, ctx->stream));
    return 0;
}

int gnn_gcn_train_step(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X, int64_t ldx, const int32_t *y, float lr,
                       float *loss_d) {
    GNN_REQUIRE(ctx && m && X && y && ldx >= m->dims[0], "gnn_gcn_train_step: bad argument");
    GNN_REQUIRE(!m->dist || ldx == m->ld[0], "gnn_gcn_train_step: row-partitioned mode needs ldx == round_up(F0,4)");
    GNN_TRY(ensure_buffers(ctx, m));
    m->last_y = y;
    m->span_used = 0;
    // Graph replay: single GPU, SGD (Adam's bias correction is a per-step kernel argument), no profiling.  Auto mode
    // turns it on for launch-bound sizes only (activations up to 32 M floats, i.e. steps of about a millisecond).
    // (`first` of the momentum update is a kernel argument too: capture only once the buffer has seen a gradient.)
    const bool want_graph = !m->dist && !m->profile && m->optimizer == 0 && m->steps >= 1 &&
                            (m->momentum == 0.f || lr == 0.f || m->vel_steps >= 1) &&
                            (m->use_graph == 1 || (m->use_graph < 0 && (int64_t)m->n_loc * m->maxld <= (1ll << 25)));
    const gnn_gcn::GraphKey key = {X, ldx, y, lr, loss_d, ctx->ws_gen};
    if (want_graph) {
        if (m->graph_exec && !(m->graph_key == key)) { // keep the previous graph as the second entry (swap)
            std::swap(m->graph_exec, m->graph_exec2);
            std::swap(m->graph_key, m->graph_key2);
            if (m->graph_exec && !(m->graph_key == key)) {
                cudaGraphExecDestroy(m->graph_exec);
                m->graph_exec = nullptr;
            }
        }
        if (!m->graph_exec) {
            const int64_t l0 = ctx->launches;
            cudaGraph_t graph = nullptr;
            GNN_CHECK_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            const int rc = train_step_body(ctx, m, X, ldx, y, lr, loss_d);
            const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
            if (rc || ce != cudaSuccess || ctx->ws_gen != key.ws_gen) {
                // a failed capture, or a workspace that grew while capturing (the eager warm-up step sizes it, so
                // this means the schedule changed): never keep a graph over stale pointers
                if (graph) cudaGraphDestroy(graph);
                if (rc) return rc;
                GNN_CHECK_CUDA(ce);
                GNN_REQUIRE(false, "gnn_gcn_train_step: workspace was reallocated during graph capture");
            }
            const cudaError_t ie = cudaGraphInstantiate(&m->graph_exec, graph, 0);
            cudaGraphDestroy(graph);
            GNN_CHECK_CUDA(ie);
            m->graph_launches = ctx->launches - l0;
            ctx->launches = l0;
            m->graph_key = key;
        }
        GNN_CHECK_CUDA(cudaGraphLaunch(m->graph_exec, ctx->stream));
        ctx->launches += m->graph_launches;
    } else {
        GNN_TRY(train_step_body(ctx, m, X, ldx, y, lr, loss_d));
    }
    m->steps++;
    if (lr != 0.f) {
        if (m->optimizer == 1) m->opt_steps++;
        else if (m->momentum != 0.f) m->vel_steps++;
    }
    if (m->profile) {
        GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < 6; i++) m->breakdown[i] = 0;
        for (size_t i = 0; i < m->span_used; i++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, m->spans[i].a, m->spans[i].b);
            m->spans[i].ms = ms;
            m->breakdown[m->spans[i].cls] += ms;
        }
    }
    return 0;
}

// staging buffers for the host-buffer entry points: two (X, y) device slots so the upload of the next step's
// inputs (copy stream) overlaps the current step's kernels (compute stream)
static int ensure_staging(gnn_ctx_t *ctx, gnn_gcn_t *m) {
    const int64_t rows_alloc = m->dist ? m->chunk : m->n_loc;
    for (int s = 0; s < 2; s++) {
        if (!m->Xs[s]) {
            GNN_CHECK_CUDA(cudaMalloc((void **)&m->Xs[s], (size_t)rows_alloc * m->ld[0] * 4));
            GNN_CHECK_CUDA(cudaMemsetAsync(m->Xs[s], 0, (size_t)rows_alloc * m->ld[0] * 4, ctx->stream));
        }
        if (!m->ys[s]) GNN_CHECK_CUDA(cudaMalloc((void **)&m->ys[s], (size_t)rows_alloc * 4));
    }
    if (!m->copy_stream) {
        GNN_CHECK_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        GNN_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_uploaded, cudaEventDisableTiming));
        GNN_CHECK_CUDA(cudaEventCreateWithFlags(&m->ev_consumed, cudaEventDisableTiming));
    }
    return 0;
}

static int upload(gnn_gcn_t *m, int slot, const float *X_h, const int32_t *y_h, cudaStream_t s) {
    const int32_t F0 = m->dims[0];
    if (m->ld[0] == F0) // dense rows: one flat DMA (a 2-D copy of 400-byte rows is an order of magnitude slower)
        GNN_CHECK_CUDA(cudaMemcpyAsync(m->Xs[slot], X_h, (size_t)m->n_loc * F0 * 4, cudaMemcpyHostToDevice, s));
    else
        GNN_CHECK_CUDA(cudaMemcpy2DAsync(m->Xs[slot], (size_t)m->ld[0] * 4, X_h, (size_t)F0 * 4, (size_t)F0 * 4,
                                         (size_t)m->n_loc, cudaMemcpyHostToDevice, s));
    GNN_CHECK_CUDA(cudaMemcpyAsync(m->ys[slot], y_h, (size_t)m->n_loc * 4, cudaMemcpyHostToDevice, s));
    return 0;
}

int gnn_gcn_prefetch_h(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X_h, const int32_t *y_h) {
    GNN_REQUIRE(ctx && m && X_h && y_h, "gnn_gcn_prefetch_h: NULL argument");
    GNN_REQUIRE(!m->prefetched, "gnn_gcn_prefetch_h: a prefetched batch is already pending");
    GNN_TRY(ensure_staging(ctx, m));
    const int slot = m->cur_slot ^ 1;
    // the back slot may still be read by the step that used it last: wait for that step on the copy stream
    GNN_CHECK_CUDA(cudaEventRecord(m->ev_consumed, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamWaitEvent(m->copy_stream, m->ev_consumed, 0));
    GNN_TRY(upload(m, slot, X_h, y_h, m->copy_stream));
    GNN_CHECK_CUDA(cudaEventRecord(m->ev_uploaded, m->copy_stream));
    m->prefetched = true;
    return 0;
}

int gnn_gcn_train_step_h(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X_h, const int32_t *y_h, float lr, float *loss_h) {
    GNN_REQUIRE(ctx && m && loss_h, "gnn_gcn_train_step_h: NULL argument");
    GNN_TRY(ensure_staging(ctx, m));
    if (m->prefetched) { // inputs were uploaded by gnn_gcn_prefetch_h: order the step after that copy
        m->cur_slot ^= 1;
        m->prefetched = false;
        GNN_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, m->ev_uploaded, 0));
        // pipelined mode: the buffers passed now are the NEXT step's inputs; their upload overlaps this step
        if (X_h && y_h) GNN_TRY(gnn_gcn_prefetch_h(ctx, m, X_h, y_h));
    } else {
        GNN_REQUIRE(X_h && y_h, "gnn_gcn_train_step_h: NULL inputs and nothing prefetched");
        GNN_TRY(upload(m, m->cur_slot, X_h, y_h, ctx->stream));
    }
    GNN_TRY(gnn_gcn_train_step(ctx, m, m->Xs[m->cur_slot], m->ld[0], m->ys[m->cur_slot], lr, m->loss_d));
    GNN_CHECK_CUDA(cudaMemcpyAsync(loss_h, m->loss_d, 4, cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int gnn_partition_panels_h(int32_t ldw, int32_t panel_cols, int32_t *c0_h, int32_t *w_h, int32_t *n_h) {
    GNN_REQUIRE(ldw > 0 && ldw % 4 == 0 && c0_h && w_h && n_h, "gnn_partition_panels_h: ldw must be a positive multiple of 4");
    const Panels P = split_panels(ldw, panel_cols);
    for (int p = 0; p < P.n; p++) { c0_h[p] = P.c0[p]; w_h[p] = P.w[p]; }
    *n_h = P.n;
    return 0;
}

int gnn_gcn_set_relu_overrides(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, const int32_t *rows, const int32_t *cols,
                               const uint8_t *positive, int64_t n) {
    GNN_REQUIRE(ctx && m && layer >= 1 && layer < m->L && n >= 0 && (n == 0 || (rows && cols && positive)),
                "gnn_gcn_set_relu_overrides: layer must be a hidden layer (1..L-1) and the arrays non-NULL");
    drop_graph(m);
    if (m->overrides.size() <= (size_t)m->L) m->overrides.resize(m->L + 1);
    m->overrides[layer].rows = rows; m->overrides[layer].cols = cols; m->overrides[layer].positive = positive;
    m->overrides[layer].n = n;
    return 0;
}

int gnn_gcn_set_train_mask(gnn_ctx_t *ctx, gnn_gcn_t *m, const uint8_t *mask, int64_t n_selected_total) {
    GNN_REQUIRE(ctx && m && (!mask || n_selected_total > 0), "invalid input, mask must be 1D and of same size with num of nodes in graph");
    drop_graph(m);
    m->train_mask = mask;
    m->n_train = n_selected_total;
    return 0;
}

int gnn_gcn_accuracy(gnn_ctx_t *ctx, gnn_gcn_t *m, const int32_t *y, const uint8_t *mask, int64_t *count) {
    GNN_REQUIRE(ctx && m && y && count, "gnn_gcn_accuracy: NULL argument");
    return gnn_argmax_correct(ctx, m->n_loc, m->dims[m->L], m->H[m->L], m->ld[m->L], y, mask, count);
}

int gnn_gcn_last_breakdown(gnn_gcn_t *m, double *ms, int n) {
    GNN_REQUIRE(m && ms, "gnn_gcn_last_breakdown: NULL argument");
    for (int i = 0; i < n && i < 6; i++) ms[i] = m->breakdown[i];
    return 0;
}

int gnn_gcn_last_spmm_spans(gnn_gcn_t *m, double *ms, double *alg_bytes, int32_t *F, int cap, int *n) {
    GNN_REQUIRE(m && n, "gnn_gcn_last_spmm_spans: NULL argument");
    int k = 0;
    for (size_t i = 0; i < m->span_used; i++) {
        if (m->spans[i].cls != CLS_SPMM) continue;
        if (k < cap) {
            if (ms) ms[k] = m->spans[i].ms;
            if (alg_bytes) alg_bytes[k] = m->spans[i].bytes;
            if (F) F[k] = m->spans[i].F;
        }
        k++;
    }
    *n = k;
    return 0;
}

int gnn_gcn_exchange_stats(const gnn_gcn_t *m, double *halo_fraction, int *halo_lists, int *split, double *interior_fraction) {
    GNN_REQUIRE(m, "gnn_gcn_exchange_stats: NULL argument");
    if (halo_fraction) *halo_fraction = m->halo_fraction;
    if (halo_lists) *halo_lists = m->halo_lists ? 1 : 0;
    if (split) *split = m->split ? 1 : 0;
    if (interior_fraction) *interior_fraction = m->grid && m->grp_rows > 0 ? (double)m->sub[0][0].n / (double)m->grp_rows : 0.0;
    return 0;
}

int gnn_gcn_exchange_mode(const gnn_gcn_t *m) {
    if (!m || !m->dist) return 0;
    if (m->grid) return 6;
    if (!m->arena) return 1;
    if (m->nccl_transport) return 4;
    return gnn_peer_arena_transport(m->arena) == 0 ? 3 : 2;
}

int gnn_gcn_spmm_stats(gnn_gcn_t *m, double *alg_bytes, int32_t *n_spmm, double *gemm_flops) {
    GNN_REQUIRE(m, "gnn_gcn_spmm_stats: NULL argument");
    if (alg_bytes) *alg_bytes = m->alg_bytes;
    if (n_spmm) *n_spmm = m->n_spmm;
    if (gemm_flops) *gemm_flops = m->gemm_flops;
    return 0;
}

} // extern "C"
