// trainer_grid.cu — the fused train step under a 2-D (row groups x feature-column groups) partition of the
// aggregation, for world sizes where the 1-D row partition is exchange-bound (8 ranks all-gather 6.05 GB
// per rank and step).
//
// world = Pr x Pc, rank = gi * Pc + gj.  Dense work (transforms, loss, bias gradients, optimiser) stays 1-D
// row-partitioned exactly as in trainer.cu: rank r owns rows [r c, (r+1) c), c = ceil(N / world), of every
// activation.  An aggregation Y = A_hat P (or A_hat^T P) of width F runs as
//   1. rows -> columns:  every rank scatters its rows of P, cut into Pc column slices (multiples of 4 columns), to
//      the Pr ranks of each column group — one kernel storing into the peers' IPC-mapped arenas (comm.cu
//      peer_scatter_kernel).  Rank (gi, gj) ends up with PC = P[all N rows, slice gj].
//      Received per rank: N F / Pc (instead of N F (P-1)/P for the all-gather of the row partition).
//   2. SpMM over the structure rows of row group gi (n_rows = Pc c, all columns) at the slice width: 1/Pr of the
//      nonzeros, 1/Pc of the columns per rank.
//   3. columns -> rows, FUSED into the aggregation kernel: the epilogue (bias / ReLU) of the SpMM stores every output
//      row straight into the arena of the rank that owns the row (YDest table, spmm.cu), at the slice's column
//      offset; a flag round inside the row group publishes completion.  Received per rank: c F (Pc-1)/Pc.
// Pr = world, Pc = 1 degenerates to the row partition (without its panel pipelining); Pr = 1, Pc = world is a pure
// feature-column partition with the whole structure on every rank.  Every output element is produced by exactly one
// rank with the single-GPU kernels, so results are bit-reproducible run to run.
//
// Ordering between steps: a region (gathered slice or output rows) is rewritten by the peers one step later; the
// gradient all-reduce of the step orders all ranks in between (forward-only calls add a one-word all-reduce).
#include "trainer.cuh"

namespace gnn {

static inline int op_of(int32_t l, int dir) { return 2 * (l - 1) + dir; }

// column slice of group j of a matrix with padded width ldF (multiple of 4): 16-byte units dealt out in order
static inline void col_slice(int32_t ldF, int Pc, int j, int32_t *c0, int32_t *w) {
    const int32_t units = ldF / 4, base = units / Pc, rem = units % Pc;
    *c0 = 4 * (j * base + (j < rem ? j : rem));
    *w = 4 * (base + (j < rem ? 1 : 0));
}
static inline int32_t max_slice(int32_t ldF, int Pc) { return 4 * (int32_t)ceil_div(ldF / 4, Pc); }

// width of the matrix aggregated by layer l (forward and backward aggregate at the same, narrower, width)
static inline int32_t agg_ld(const gnn_gcn *m, int32_t l) { return m->agg_first[l] ? m->ld[l - 1] : m->ld[l]; }
static inline int32_t agg_f(const gnn_gcn *m, int32_t l) { return m->agg_first[l] ? m->dims[l - 1] : m->dims[l]; }

static inline float *y_region(gnn_gcn *m, int op) {
    return reinterpret_cast<float *>(peer_base(m->arena, m->gi * m->Pc + m->gj) + m->y_off[op]);
}
float *dz_buffer_grid(gnn_gcn *m, int32_t l) { return ((m->L - l) & 1) ? m->G1 : m->G0; }

void recompute_stats_grid(gnn_gcn *m) {
    const gnn_graph *g = m->g;
    m->alg_bytes = 0; m->gemm_flops = 0; m->n_spmm = 0;
    for (int32_t l = 1; l <= m->L; l++) {
        int32_t c0, w;
        col_slice(agg_ld(m, l), m->Pc, m->gj, &c0, &w);
        const int32_t f = agg_f(m, l) - c0 < w ? agg_f(m, l) - c0 : w;
        if (f > 0) {
            m->alg_bytes += spmm_alg_bytes(m->grp_rows, g->nnz, f); m->n_spmm++;
            if (!(m->agg_first[l] && l == 1)) { m->alg_bytes += spmm_alg_bytes(m->grp_rows, g->nnz_t, f); m->n_spmm++; }
        }
        m->gemm_flops += 2.0 * m->n_loc * m->dims[l - 1] * m->dims[l] * (l > 1 ? 3 : 2);
    }
}


// ---- halo-only exchange and interior/boundary split: set-up kernels ----------------------------------------------
__global__ void need_mark_kernel(const int32_t *__restrict__ idx, int64_t nnz, uint8_t *__restrict__ need) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) need[idx[k]] = 1; // benign race: every writer stores 1
}
__global__ void need_flags_kernel(const uint8_t *__restrict__ need, int64_t n, uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) flags[i] = (i < n && need[i]) ? 1u : 0u;
}
__global__ void list_compact_kernel(const uint8_t *__restrict__ need, const uint32_t *__restrict__ pos, int64_t n,
                                    int32_t *__restrict__ list) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && need[i]) list[pos[i]] = (int32_t)i;
}
// one warp per structure row: interior iff every column is one of the rank's own rows [own_lo, own_hi)
__global__ void interior_flag_kernel(const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx, int32_t n_rows,
                                     int32_t own_lo, int32_t own_hi, uint32_t *__restrict__ iflag, uint32_t *__restrict__ bflag) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row > n_rows) return;
    if (row == n_rows) { if (lane == 0) { iflag[row] = 0; bflag[row] = 0; } return; }
    bool in = true;
    for (int32_t k = ptr[row] + lane; k < ptr[row + 1]; k += 32) {
        const int32_t c = idx[k];
        if (c < own_lo || c >= own_hi) in = false;
    }
    in = __all_sync(0xffffffffu, in);
    if (lane == 0) { iflag[row] = in ? 1u : 0u; bflag[row] = in ? 0u : 1u; }
}
__global__ void sub_rows_kernel(const uint32_t *__restrict__ sel, const uint32_t *__restrict__ pos, const int32_t *__restrict__ ptr,
                                int32_t n_rows, int32_t *__restrict__ rows, uint32_t *__restrict__ len) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n_rows && sel[r]) { rows[pos[r]] = (int32_t)r; len[pos[r]] = (uint32_t)(ptr[r + 1] - ptr[r]); }
}
__global__ void sub_copy_kernel(const int32_t *__restrict__ rows, const int32_t *__restrict__ sub_ptr, int32_t n_sub,
                                const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx, const float *__restrict__ val,
                                int32_t *__restrict__ sub_idx, float *__restrict__ sub_val) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n_sub) return;
    const int32_t r = rows[i], s0 = ptr[r], n = ptr[r + 1] - s0, d0 = sub_ptr[i];
    for (int32_t k = lane; k < n; k += 32) { sub_idx[d0 + k] = idx[s0 + k]; sub_val[d0 + k] = val[s0 + k]; }
}
static inline unsigned g256(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, 256); }

// CSR of the selected rows (sel[r] = 1) of a structure block, with the list of the original row ids
static int build_sub_csr(gnn_ctx *ctx, const uint32_t *sel, const int32_t *ptr, const int32_t *idx, const float *val,
                         int32_t n_rows, int32_t parent_max, gnn_gcn::SubCsr *out) {
    cudaStream_t s = ctx->stream;
    uint32_t *pos = nullptr, *len = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&pos, (size_t)(n_rows + 1) * 4, s));
    GNN_TRY(exclusive_scan_u32(ctx, sel, pos, n_rows + 1, nullptr));
    uint32_t n_sub = 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(&n_sub, pos + n_rows, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    out->n = (int32_t)n_sub;
    out->max_nnz = parent_max;
    GNN_CHECK_CUDA(cudaMalloc((void **)&out->rows, (size_t)(n_sub + 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&out->ptr, (size_t)(n_sub + 2) * 4));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&len, (size_t)(n_sub + 2) * 4, s));
    GNN_CHECK_CUDA(cudaMemsetAsync(len, 0, (size_t)(n_sub + 2) * 4, s));
    sub_rows_kernel<<<g256(n_rows), 256, 0, s>>>(sel, pos, ptr, n_rows, out->rows, len);
    GNN_LAUNCHED(ctx);
    GNN_TRY(exclusive_scan_u32(ctx, len, reinterpret_cast<uint32_t *>(out->ptr), (int64_t)n_sub + 1, nullptr));
    int32_t nnz = 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(&nnz, out->ptr + n_sub, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    out->nnz = nnz;
    GNN_CHECK_CUDA(cudaMalloc((void **)&out->idx, (size_t)(nnz ? nnz : 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&out->val, (size_t)(nnz ? nnz : 1) * 4));
    sub_copy_kernel<<<g256((int64_t)n_sub * 32), 256, 0, s>>>(out->rows, out->ptr, (int32_t)n_sub, ptr, idx, val, out->idx, out->val);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaFreeAsync(pos, s));
    GNN_CHECK_CUDA(cudaFreeAsync(len, s));
    return 0;
}

// Collective set-up after the arena exists.  (1) every rank publishes which node rows its structure block reads (forward
// and backward blocks), (2) reads the peers' flags for its OWN rows and keeps, per destination, the list of rows that
// destination needs (the halo-only exchange; skipped when the peers need practically everything, as on the synthetic
// power-law graphs), (3) splits its structure rows into interior (only own rows needed) and boundary rows.
static int setup_halo(gnn_ctx *ctx, gnn_gcn *m) {
    const gnn_graph *g = m->g;
    cudaStream_t s = ctx->stream;
    const int64_t N = m->n_glob;
    const int64_t own_lo = std::min<int64_t>(N, (int64_t)ctx->rank * m->chunk), own_hi = own_lo + m->n_loc;
    const size_t need_stride = (size_t)round_up(N, 256);
    double listed = 0, possible = 0;
    for (int dir = 0; dir < 2; dir++) {
        const int32_t *idx = dir ? g->rowidx : g->colidx;
        const int64_t nnz = dir ? g->nnz_t : g->nnz;
        uint8_t *need = reinterpret_cast<uint8_t *>(peer_base(m->arena, ctx->rank) + m->need_off + dir * need_stride);
        GNN_CHECK_CUDA(cudaMemsetAsync(need, 0, need_stride, s));
        need_mark_kernel<<<g256(nnz), 256, 0, s>>>(idx, nnz, need);
        GNN_LAUNCHED(ctx);
    }
    GNN_TRY(gnn_allreduce_sum(ctx, m->grads + m->n_params + 1, 1)); // every rank's flags are written
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    uint32_t *flags = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&flags, (size_t)(m->n_loc + 2) * 4, s));
    for (int dir = 0; dir < 2; dir++) {
        m->send_list[dir].assign(ctx->world, nullptr);
        m->send_cnt[dir].assign(ctx->world, 0);
        for (int q = 0; q < ctx->world; q++) {
            if (q == ctx->rank || m->n_loc == 0) continue;
            const uint8_t *need_q = reinterpret_cast<const uint8_t *>(peer_base(m->arena, q) + m->need_off + dir * need_stride) + own_lo;
            need_flags_kernel<<<g256(m->n_loc + 1), 256, 0, s>>>(need_q, m->n_loc, flags);
            GNN_LAUNCHED(ctx);
            GNN_TRY(exclusive_scan_u32(ctx, flags, flags, m->n_loc + 1, nullptr));
            uint32_t cnt = 0;
            GNN_CHECK_CUDA(cudaMemcpyAsync(&cnt, flags + m->n_loc, 4, cudaMemcpyDeviceToHost, s));
            GNN_CHECK_CUDA(cudaStreamSynchronize(s));
            GNN_CHECK_CUDA(cudaMalloc((void **)&m->send_list[dir][q], (size_t)(cnt ? cnt : 1) * 4));
            list_compact_kernel<<<g256(m->n_loc), 256, 0, s>>>(need_q, flags, m->n_loc, m->send_list[dir][q]);
            GNN_LAUNCHED(ctx);
            m->send_cnt[dir][q] = cnt;
            listed += cnt;
            possible += (double)m->n_loc;
        }
    }
    GNN_CHECK_CUDA(cudaFreeAsync(flags, s));
    GNN_TRY(gnn_allreduce_sum(ctx, m->grads + m->n_params + 1, 1)); // nobody is still reading this rank's flags
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    m->halo_fraction = possible > 0 ? listed / possible : 1.0;
    m->halo_lists = m->halo_fraction < 0.9;
    if (const char *e = getenv("GNN_HALO")) m->halo_lists = atoi(e) != 0; // ablation: 0 = always send every row
    // interior / boundary split of the structure block
    uint32_t *iflag = nullptr, *bflag = nullptr;
    const int32_t rows_max = g->n_rows > g->t_rows ? g->n_rows : g->t_rows;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&iflag, (size_t)(rows_max + 1) * 4, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&bflag, (size_t)(rows_max + 1) * 4, s));
    bool want_split = true;
    for (int dir = 0; dir < 2 && want_split; dir++) {
        const int32_t *ptr = dir ? g->colptr : g->rowptr, *idx = dir ? g->rowidx : g->colidx;
        const float *val = dir ? g->valT : g->val;
        const int32_t n_rows = dir ? g->t_rows : g->n_rows;
        interior_flag_kernel<<<g256((int64_t)(n_rows + 1) * 32), 256, 0, s>>>(ptr, idx, n_rows, (int32_t)own_lo, (int32_t)own_hi, iflag, bflag);
        GNN_LAUNCHED(ctx);
        GNN_TRY(build_sub_csr(ctx, iflag, ptr, idx, val, n_rows, dir ? g->max_col_nnz : g->max_row_nnz, &m->sub[dir][0]));
        // worth two launches only when a real share of the rows can start early
        if (m->sub[dir][0].n < n_rows / 10) { want_split = false; break; }
        GNN_TRY(build_sub_csr(ctx, bflag, ptr, idx, val, n_rows, dir ? g->max_col_nnz : g->max_row_nnz, &m->sub[dir][1]));
    }
    GNN_CHECK_CUDA(cudaFreeAsync(iflag, s));
    GNN_CHECK_CUDA(cudaFreeAsync(bflag, s));
    m->split = want_split;
    if (const char *e = getenv("GNN_SPLIT")) m->split = m->split && atoi(e) != 0; // ablation: 0 = one launch over all rows
    // ranks may decide differently on lists (a sender-side choice) but the split changes nothing a peer can see either
    return 0;
}

// step 1: scatter the rank's rows of src[n_loc, ldF] to every rank (each receives its column group's slice)
static int issue_scatter(gnn_ctx *ctx, gnn_gcn *m, int op, const float *src, int64_t ld_src, int32_t ldF) {
    if (m->scattered[op]) return 0;
    Prof pr(ctx, m, CLS_OTHER);
    const int dir = op & 1;
    int ranks[16];
    size_t offs[16];
    int32_t c0s[16], ws[16];
    const int32_t *lists[16];
    int64_t cnts[16];
    int n = 0;
    for (int i = 1; i <= ctx->world; i++) { // staggered start; the own rank last
        const int q = (ctx->rank + i) % ctx->world;
        int32_t c0, w;
        col_slice(ldF, m->Pc, q % m->Pc, &c0, &w);
        // a column group that gets no columns of a narrow matrix (w == 0) still receives the flag: it waits for every rank
        if (q == ctx->rank && m->split) { // the own slice is copied on the compute stream: the interior rows start on it at once
            if (w > 0 && m->n_loc > 0) {
                float *own = reinterpret_cast<float *>(peer_base(m->arena, q) + m->pc_off[op]) + (size_t)ctx->rank * m->chunk * w;
                GNN_TRY(copy2d(ctx, own, w, src + c0, ld_src, m->n_loc, w));
            }
            continue;
        }
        ranks[n] = q;
        offs[n] = m->pc_off[op] + (size_t)ctx->rank * m->chunk * w * 4;
        c0s[n] = c0;
        ws[n] = w;
        lists[n] = (m->halo_lists && q != ctx->rank) ? m->send_list[dir][q] : nullptr; // only the rows q's structure reads
        cnts[n] = lists[n] ? m->send_cnt[dir][q] : m->n_loc;
        n++;
    }
    GNN_TRY(peer_scatter_begin(ctx, m->arena, 2 * op, src, ld_src, m->n_loc, n, ranks, offs, c0s, ws, lists, cnts));
    m->scattered[op] = 1;
    return 0;
}

// steps 1-3 for one aggregation; the result rows of this rank are in y_region(op) [n_loc, ldF] afterwards
static int aggregate_grid(gnn_ctx *ctx, gnn_gcn *m, int op, int transpose, const float *src, int64_t ld_src, int32_t F,
                          int32_t ldF, const float *bias, int relu) {
    const gnn_graph *g = m->g;
    GNN_TRY(issue_scatter(ctx, m, op, src, ld_src, ldF));
    m->scattered[op] = 0; // consumed: the next step scatters again
    const uint32_t all = ctx->world >= 32 ? 0xffffffffu : ((1u << ctx->world) - 1u);
    uint32_t rowgrp = 0;
    for (int q = 0; q < m->Pc; q++) rowgrp |= 1u << (m->gi * m->Pc + q);
    int32_t c0, w;
    col_slice(ldF, m->Pc, m->gj, &c0, &w);
    const int32_t f = F - c0 < w ? F - c0 : w;
    const bool work = w > 0 && f > 0 && m->grp_rows > 0;
    YDest d;
    d.row_map = nullptr;
    d.rows_per = (int32_t)m->chunk;
    for (int q = 0; q < SPMM_MAX_DEST; q++)
        d.base[q] = q < m->Pc ? reinterpret_cast<float *>(peer_base(m->arena, m->gi * m->Pc + q) + m->y_off[op]) + c0 : nullptr;
    const float *PC = reinterpret_cast<const float *>(peer_base(m->arena, ctx->rank) + m->pc_off[op]);
    const float *bs = bias ? bias + c0 : nullptr;
    auto sub_spmm = [&](const gnn_gcn::SubCsr &sc) -> int { // aggregation over a row subset, rows routed through row_map
        if (!work || sc.n == 0) return 0;
        Prof pr(ctx, m, CLS_SPMM, f, spmm_alg_bytes(sc.n, sc.nnz, f));
        YDest ds = d;
        ds.row_map = sc.rows;
        return spmm_launch(ctx, sc.n, 0, sc.nnz, sc.ptr, sc.idx, sc.val, 1, sc.max_nnz, PC, w, f, nullptr, ldF, bs, relu, nullptr, 0,
                           true, &ds);
    };
    if (m->split) GNN_TRY(sub_spmm(m->sub[transpose ? 1 : 0][0])); // interior rows: only the rank's own rows are read
    {
        Prof pr(ctx, m, CLS_OTHER);
        // also ordered after this rank's own scatter kernel (it reads `src`, which later kernels overwrite); with the
        // split the interior aggregation is already enqueued, so nothing is lost by waiting here
        GNN_TRY(peer_wait_mask(ctx, m->arena, 2 * op, all, true));
    }
    if (m->split) {
        GNN_TRY(sub_spmm(m->sub[transpose ? 1 : 0][1]));           // boundary rows: after every peer's rows have landed
    } else if (work) {
        const int64_t nnz = transpose ? g->nnz_t : g->nnz;
        Prof pr(ctx, m, CLS_SPMM, f, spmm_alg_bytes(m->grp_rows, nnz, f));
        GNN_TRY(spmm_rows_range(ctx, g, transpose, 0, (int32_t)m->grp_rows, 0, nnz, PC, w, f, nullptr, ldF, bs, relu, nullptr, 0, &d));
    }
    Prof pr(ctx, m, CLS_OTHER);
    GNN_TRY(peer_signal(ctx, m->arena, 2 * op + 1, rowgrp));
    GNN_TRY(peer_wait_mask(ctx, m->arena, 2 * op + 1, rowgrp, false));
    return 0;
}

int forward_grid(gnn_ctx *ctx, gnn_gcn *m, const float *X, int64_t ldx) {
    const float *Hin = X;
    int64_t ld_in = ldx;
    for (int32_t l = 1; l <= m->L; l++) {
        const int32_t Fi = m->dims[l - 1], Fo = m->dims[l];
        const float *W = m->params + m->w_off[l], *b = m->params + m->b_off[l];
        const int relu = l < m->L;
        const int op = op_of(l, 0);
        if (m->agg_first[l]) {
            GNN_TRY(aggregate_grid(ctx, m, op, 0, Hin, ld_in, Fi, m->ld[l - 1], nullptr, 0)); // M_l = A_hat H_{l-1}
            if (m->n_loc > 0) {
                Prof pr(ctx, m, CLS_GEMM);
                GNN_TRY(gnn_gemm_nt(ctx, m->n_loc, Fo, Fi, m->M[l], m->ld[l - 1], W, Fi, m->H[l], m->ld[l], b, relu, m->precision));
            }
        } else {
            if (m->n_loc > 0) {
                Prof pr(ctx, m, CLS_GEMM);
                GNN_TRY(gnn_gemm_nt(ctx, m->n_loc, Fo, Fi, Hin, ld_in, W, Fi, m->S1, m->ld[l], nullptr, 0, m->precision));
            }
            GNN_TRY(aggregate_grid(ctx, m, op, 0, m->S1, m->ld[l], Fo, m->ld[l], b, relu)); // H_l = act(A_hat P + b)
        }
        GNN_TRY(apply_relu_overrides(ctx, m, l, 0, m->n_loc));
        Hin = m->H[l];
        ld_in = m->ld[l];
    }
    return 0;
}

// dZ_L has been written to dz_buffer_grid(m, L)
int backward_grid(gnn_ctx *ctx, gnn_gcn *m, const float *X, int64_t ldx) {
    for (int32_t l = m->L; l >= 1; l--) {
        const int32_t Fi = m->dims[l - 1], Fo = m->dims[l];
        const float *W = m->params + m->w_off[l];
        float *dW = m->grads + m->w_off[l], *db = m->grads + m->b_off[l];
        const float *Hin = l > 1 ? m->H[l - 1] : X;
        const int64_t ld_in = l > 1 ? m->ld[l - 1] : ldx;
        float *dz = dz_buffer_grid(m, l);
        const int op = op_of(l, 1);
        if (l < m->L && m->n_loc > 0) { // db_L comes out of the loss kernel; runs while a pre-issued scatter is in flight
            Prof pr(ctx, m, CLS_BIAS);
            GNN_TRY(colsum(ctx, m->n_loc, Fo, dz, m->ld[l], db));
        } else if (l < m->L) {
            GNN_CHECK_CUDA(cudaMemsetAsync(db, 0, (size_t)Fo * 4, ctx->stream));
        }
        if (m->agg_first[l]) {
            if (l > 1) { // dM = dZ W, exchanged while dW is computed
                if (m->n_loc > 0) {
                    Prof pr(ctx, m, CLS_GEMM);
                    GNN_TRY(gnn_gemm_nn(ctx, m->n_loc, Fi, Fo, dz, m->ld[l], W, Fi, m->S1, m->ld[l - 1], nullptr, 0, m->precision));
                }
                GNN_TRY(issue_scatter(ctx, m, op, m->S1, m->ld[l - 1], m->ld[l - 1]));
            }
            if (m->n_loc > 0) {
                Prof pr(ctx, m, CLS_GEMM);
                GNN_TRY(gnn_gemm_tn(ctx, m->n_loc, Fo, Fi, dz, m->ld[l], m->M[l], m->ld[l - 1], dW, Fi, m->precision));
            } else {
                GNN_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)Fo * Fi * 4, ctx->stream));
            }
            if (l > 1) { // dZ_{l-1} = (A_hat^T dM) . [H_{l-1} > 0]: the mask rows live with the row's owner
                GNN_TRY(aggregate_grid(ctx, m, op, 1, m->S1, m->ld[l - 1], Fi, m->ld[l - 1], nullptr, 0));
                if (m->n_loc > 0)
                    GNN_TRY(gnn_relu_bwd(ctx, m->n_loc, Fi, y_region(m, op), m->ld[l - 1], Hin, ld_in, dz_buffer_grid(m, l - 1), m->ld[l - 1]));
            }
        } else {
            GNN_TRY(aggregate_grid(ctx, m, op, 1, dz, m->ld[l], Fo, m->ld[l], nullptr, 0)); // dP = A_hat^T dZ
            const float *dP = y_region(m, op);
            if (l > 1) {
                float *dn = dz_buffer_grid(m, l - 1);
                if (m->n_loc > 0) {
                    Prof pr(ctx, m, CLS_GEMM);
                    GNN_TRY(gnn_gemm_nn(ctx, m->n_loc, Fi, Fo, dP, m->ld[l], W, Fi, dn, m->ld[l - 1], Hin, ld_in, m->precision));
                }
                // the next aggregation input is ready: start its exchange now, the dW product below overlaps it
                if (!m->agg_first[l - 1]) GNN_TRY(issue_scatter(ctx, m, op_of(l - 1, 1), dn, m->ld[l - 1], m->ld[l - 1]));
            }
            if (m->n_loc > 0) {
                Prof pr(ctx, m, CLS_GEMM);
                GNN_TRY(gnn_gemm_tn(ctx, m->n_loc, Fo, Fi, dP, m->ld[l], Hin, ld_in, dW, Fi, m->precision));
            } else {
                GNN_CHECK_CUDA(cudaMemsetAsync(dW, 0, (size_t)Fo * Fi * 4, ctx->stream));
            }
        }
    }
    return 0;
}

} // namespace gnn

using namespace gnn;

extern "C" {

static int create_grid_impl(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t L, const int32_t *dims, int32_t Pr, int32_t Pc, gnn_gcn *m);

int gnn_gcn_create_grid(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t L, const int32_t *dims, int32_t Pr, int32_t Pc,
                        gnn_gcn_t **out) {
    GNN_REQUIRE(ctx && g && dims && out && L >= 1, "gnn_gcn_create_grid: bad argument");
    GNN_REQUIRE(Pr >= 1 && Pc >= 1 && Pr * Pc == ctx->world && ctx->world > 1 && ctx->world <= 16 && Pc <= SPMM_MAX_DEST,
                "gnn_gcn_create_grid: Pr x Pc = %d x %d does not match the communicator (world %d, Pc <= %d)", Pr, Pc,
                ctx->world, SPMM_MAX_DEST);
    GNN_REQUIRE(g->val && g->colptr && g->valT, "gnn_gcn_create_grid: graph must be a gnn_graph_slice_rows slice of a normalised graph with its CSC");
    for (int32_t l = 0; l <= L; l++) GNN_REQUIRE(dims[l] > 0, "dims cannot be empty or zero");
    const int64_t N = g->n_cols, c = ceil_div(N, ctx->world);
    const int gi = ctx->rank / Pc, gj = ctx->rank % Pc;
    const int64_t glo = std::min<int64_t>(N, (int64_t)gi * Pc * c), ghi = std::min<int64_t>(N, (int64_t)(gi + 1) * Pc * c);
    GNN_REQUIRE(g->n_rows == ghi - glo, "gnn_gcn_create_grid: the structure must hold rows [%lld, %lld) of the graph (row group %d), it has %d rows",
                (long long)glo, (long long)ghi, gi, g->n_rows);
    gnn_gcn *m = new gnn_gcn();
    m->g = g; m->L = L; m->grid = true; m->dist = true;
    m->Pr = Pr; m->Pc = Pc; m->gi = gi; m->gj = gj;
    m->n_glob = N; m->chunk = c; m->grp_rows = g->n_rows;
    const int rc = create_grid_impl(ctx, g, L, dims, Pr, Pc, m);
    if (rc) { // nothing half-built survives a failed set-up (every rank fails together: the set-up is collective)
        const std::string msg = gnn_last_error();
        gnn_gcn_destroy(ctx, m);
        set_error("%s", msg.c_str());
        return rc;
    }
    *out = m;
    return 0;
}

static int create_grid_impl(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t L, const int32_t *dims, int32_t Pr, int32_t Pc, gnn_gcn *m) {
    const int64_t N = g->n_cols, c = m->chunk;
    m->n_glob = N; m->chunk = c; m->grp_rows = g->n_rows;
    m->n_loc = std::max<int64_t>(0, std::min<int64_t>(N, (int64_t)(ctx->rank + 1) * c) - std::min<int64_t>(N, (int64_t)ctx->rank * c));
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
        off = round_up(off, 4);
    }
    m->n_params = off;
    auto alloc = [&](float **p, int64_t n) -> int {
        GNN_CHECK_CUDA(cudaMalloc((void **)p, (size_t)(n > 0 ? n : 1) * 4));
        GNN_CHECK_CUDA(cudaMemsetAsync(*p, 0, (size_t)(n > 0 ? n : 1) * 4, ctx->stream));
        m->owned.push_back(*p);
        return 0;
    };
    GNN_TRY(alloc(&m->params, m->n_params));
    GNN_TRY(alloc(&m->grads, m->n_params + 4));
    GNN_TRY(alloc(&m->S1, c * m->maxld));
    GNN_TRY(alloc(&m->G0, c * m->maxld));
    GNN_TRY(alloc(&m->G1, c * m->maxld));
    GNN_TRY(alloc(&m->loss_d, 4));
    // arena: per aggregation op one gathered-slice region [world c, widest slice] and one output region [c, width];
    // offsets are the same on every rank (remote addressing), so they are sized for the widest column group
    m->pc_off.assign(2 * L, 0);
    m->y_off.assign(2 * L, 0);
    m->scattered.assign(2 * L, 0);
    size_t aoff = 0;
    for (int32_t l = 1; l <= L; l++)
        for (int dir = 0; dir < 2; dir++) {
            const int op = op_of(l, dir);
            // (sized for either layer order: the "agg_first_mask" option may flip a layer later)
            const int32_t wide = m->ld[l - 1] > m->ld[l] ? m->ld[l - 1] : m->ld[l];
            m->pc_off[op] = aoff;
            aoff += (size_t)round_up((int64_t)ctx->world * c * max_slice(wide, Pc) * 4, 256);
            m->y_off[op] = aoff;
            aoff += (size_t)round_up(c * (int64_t)wide * 4, 256);
        }
    m->need_off = aoff;
    aoff += 2 * (size_t)round_up(N, 256); // which node rows this rank's forward / backward structure block reads (setup_halo)
    GNN_TRY(gnn_peer_arena_create(ctx, aoff, &m->arena)); // no NCCL variant of this mode: on failure the caller picks the row partition
    m->H.assign(L + 1, nullptr);
    m->M.assign(L + 1, nullptr);
    for (int32_t l = 1; l <= L; l++) {
        // both layer orders stay available ("agg_first_mask"): M_l and the transform-first H_l alias the forward
        // aggregation's output region, the aggregate-first H_l is a local buffer
        m->M[l] = y_region(m, op_of(l, 0));
    }
    m->H_local.assign(L + 1, nullptr);
    for (int32_t l = 1; l <= L; l++) GNN_TRY(alloc(&m->H_local[l], c * m->ld[l]));
    for (int32_t l = 1; l <= L; l++) m->H[l] = m->agg_first[l] ? m->H_local[l] : y_region(m, op_of(l, 0));
    GNN_TRY(setup_halo(ctx, m));
    recompute_stats_grid(m);
    return 0;
}


int gnn_partition_col_slice_h(int32_t ldw, int32_t Pc, int32_t j, int32_t *c0_h, int32_t *w_h) {
    GNN_REQUIRE(ldw > 0 && ldw % 4 == 0 && Pc >= 1 && j >= 0 && j < Pc && c0_h && w_h,
                "gnn_partition_col_slice_h: ldw must be a positive multiple of 4 and 0 <= j < Pc");
    col_slice(ldw, Pc, j, c0_h, w_h);
    return 0;
}

int gnn_partition_grid_h(int64_t N, int32_t world, int32_t Pc, int32_t rank, int64_t *rows_lo_h, int64_t *rows_hi_h,
                         int64_t *group_lo_h, int64_t *group_hi_h) {
    GNN_REQUIRE(N > 0 && world >= 1 && Pc >= 1 && world % Pc == 0 && rank >= 0 && rank < world,
                "gnn_partition_grid_h: world must be a multiple of Pc and 0 <= rank < world");
    const int64_t c = ceil_div(N, world);
    const int64_t gi = rank / Pc;
    if (rows_lo_h) *rows_lo_h = std::min<int64_t>(N, rank * c);
    if (rows_hi_h) *rows_hi_h = std::min<int64_t>(N, (rank + 1) * c);
    if (group_lo_h) *group_lo_h = std::min<int64_t>(N, gi * Pc * c);
    if (group_hi_h) *group_hi_h = std::min<int64_t>(N, (gi + 1) * Pc * c);
    return 0;
}

} // extern "C"
