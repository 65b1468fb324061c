// graph.cu — device structure build (K1-K3): COO -> CSR (sorted, de-duplicated, diagonal policy),
// CSR -> CSC + permutation, degree / D^-1/2 / edge values.  Integer results are bit-exact against the
// CPU oracle, which is itself pinned to the reference's dense round trip (src/graph.cpp:21-75).
#include "common.cuh"

namespace gnn {

static int bits_for(int64_t n) { // bits to represent values in [0, n)
    int b = 1;
    while (((int64_t)1 << b) < n) b++;
    return b;
}

__global__ void make_keys_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t E,
                                 int32_t n_rows, int32_t n_cols, int cb, int add_diag, uint64_t *__restrict__ keys,
                                 int *__restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) {
        const int32_t r = src[i], c = dst[i];
        if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) {
            *err = 1;
            keys[i] = 0;
        } else {
            keys[i] = ((uint64_t)(uint32_t)r << cb) | (uint32_t)c;
        }
    } else if (add_diag && i < E + n_rows) {
        const uint64_t d = (uint64_t)(i - E);
        keys[i] = (d << cb) | d;
    }
}

__global__ void unique_flags_kernel(const uint64_t *__restrict__ keys, int64_t m, int cb, int drop_diag,
                                    uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t k = keys[i];
    bool keep = (i == 0) || (k != keys[i - 1]);
    if (drop_diag && (k >> cb) == (k & (((uint64_t)1 << cb) - 1))) keep = false;
    flags[i] = keep ? 1u : 0u;
}

// pos has m+1 entries (exclusive scan incl. total): element i was kept iff pos[i+1] != pos[i]
__global__ void compact_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pos, int64_t m, int cb,
                               uint64_t *__restrict__ ukeys, int32_t *__restrict__ colidx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t p = pos[i];
    if (pos[i + 1] != p) {
        const uint64_t k = keys[i];
        ukeys[p] = k;
        colidx[p] = (int32_t)(k & (((uint64_t)1 << cb) - 1));
    }
}

// ptr[r] = first position in sorted `ukeys` whose (key >> shift) & mask... >= r  (lower bound)
__global__ void lower_bound_kernel(const uint64_t *__restrict__ ukeys, int64_t nnz, int shift, uint64_t field_mask,
                                   int32_t n, int32_t *__restrict__ ptr) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    int64_t lo = 0, hi = nnz;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const uint64_t f = (ukeys[mid] >> shift) & field_mask;
        if (f < (uint64_t)r) lo = mid + 1;
        else hi = mid;
    }
    ptr[r] = (int32_t)lo;
}

// out[0] = max row length, out[1] = min row length (out[1] preset to INT_MAX)
__global__ void max_diff_kernel(const int32_t *__restrict__ ptr, int32_t n, int32_t *__restrict__ out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int32_t d = (r < n) ? ptr[r + 1] - ptr[r] : 0;
    int32_t m = (r < n) ? d : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d = max(d, __shfl_xor_sync(0xffffffffu, d, o));
        m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (d > 0) atomicMax(out, d);
        atomicMin(out + 1, m);
    }
}

// one warp per row: keys[k] = row << cb | col, vals[k] = k
__global__ void csr_to_keys_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                   int32_t n_rows, int cb, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const int32_t b = rowptr[w], e = rowptr[w + 1];
    for (int32_t k = b + lane; k < e; k += 32) {
        keys[k] = ((uint64_t)(uint32_t)w << cb) | (uint32_t)colidx[k];
        vals[k] = (uint32_t)k;
    }
}

__global__ void unpack_csc_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t nnz,
                                  int cb, int32_t *__restrict__ rowidx, int32_t *__restrict__ perm) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    rowidx[i] = (int32_t)(keys[i] >> cb);
    perm[i] = (int32_t)vals[i];
}

__global__ void arrays_equal_kernel(const int32_t *__restrict__ a, const int32_t *__restrict__ b, int64_t n,
                                    int *__restrict__ differ) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) *differ = 1;
}

__global__ void degree_kernel(const int32_t *__restrict__ rowptr, int32_t n, int32_t *__restrict__ deg,
                              float *__restrict__ dinv) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int32_t d = rowptr[r + 1] - rowptr[r];
    deg[r] = d;
    // deg^-1/2, correctly rounded from fp64 (the reference's std::pow(float,-0.5f) is within 1 ulp of this)
    dinv[r] = d > 0 ? (float)(1.0 / sqrt((double)d)) : 0.0f;
}

// one warp per row: val[k] = dinv[row] * dinv[col]
__global__ void edge_val_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                const float *__restrict__ dinv, int32_t n_rows, float *__restrict__ val) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const int32_t b = rowptr[w], e = rowptr[w + 1];
    const float dr = dinv[w];
    for (int32_t k = b + lane; k < e; k += 32) val[k] = dr * dinv[colidx[k]];
}

// ---- weighted adjacency (edge_attr): keys carry the edge position so that, after a STABLE sort, the last element of
// a run of equal (row, col) keys is the last write of the reference's assignment loop (src/graph.cpp:35-40)
__global__ void make_keys_w_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t E,
                                   int32_t n, int cb, int add_diag, uint64_t *__restrict__ keys, uint32_t *__restrict__ pos,
                                   int *__restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < E) {
        const int32_t r = src[i], c = dst[i];
        if (r < 0 || r >= n || c < 0 || c >= n) {
            *err = 1;
            keys[i] = 0;
        } else {
            keys[i] = ((uint64_t)(uint32_t)r << cb) | (uint32_t)c;
        }
        pos[i] = (uint32_t)i;
    } else if (add_diag && i < E + n) { // appended after every edge: the forced diagonal (fill_diagonal_(1)) wins
        const uint64_t d = (uint64_t)(i - E);
        keys[i] = (d << cb) | d;
        pos[i] = (uint32_t)i;
    }
}
__global__ void last_flags_kernel(const uint64_t *__restrict__ keys, int64_t m, uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    flags[i] = (i == m - 1 || keys[i] != keys[i + 1]) ? 1u : 0u;
}
__global__ void compact_w_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ pos_in,
                                 const uint32_t *__restrict__ pos, int64_t m, int cb, int64_t E,
                                 const float *__restrict__ w, uint64_t *__restrict__ ukeys, int32_t *__restrict__ colidx,
                                 float *__restrict__ val0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t p = pos[i];
    if (pos[i + 1] != p) {
        const uint64_t k = keys[i];
        ukeys[p] = k;
        colidx[p] = (int32_t)(k & (((uint64_t)1 << cb) - 1));
        const uint32_t e = pos_in[i];
        val0[p] = e < E ? w[e] : 1.0f;
    }
}
// weighted degree = row sum of the raw weights (one warp per row), dinv = deg^-1/2, val = (w * dinv[r]) * dinv[c]
__global__ void degree_w_kernel(const int32_t *__restrict__ rowptr, const float *__restrict__ val0, int32_t n_rows,
                                int32_t *__restrict__ deg, float *__restrict__ dinv) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    float s = 0.f;
    for (int32_t k = rowptr[w] + lane; k < rowptr[w + 1]; k += 32) s += val0[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        deg[w] = rowptr[w + 1] - rowptr[w];
        dinv[w] = (float)(1.0 / sqrt((double)s));
    }
}
__global__ void edge_val_w_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                  const float *__restrict__ val0, const float *__restrict__ dinv, int32_t n_rows,
                                  float *__restrict__ val) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const float dr = dinv[w];
    for (int32_t k = rowptr[w] + lane; k < rowptr[w + 1]; k += 32) val[k] = (val0[k] * dr) * dinv[colidx[k]];
}

// graph::GCNConv::forward as written (reference src/graph.cpp:176-185) on the loop-free adjacency A0:
//   dinv = (rowsum(A0) + 1)^-1/2 ;  norm[r] = dinv[r] * sum_{c in row r} dinv[c]     (one warp per row, fixed order)
__global__ void aswritten_dinv_kernel(const int32_t *__restrict__ rowptr, int32_t n, int32_t *__restrict__ deg,
                                      float *__restrict__ dinv) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int32_t d = rowptr[r + 1] - rowptr[r] + 1;
    deg[r] = d;
    dinv[r] = (float)(1.0 / sqrt((double)d));
}
__global__ void aswritten_norm_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                      const float *__restrict__ dinv, int32_t n_rows, float *__restrict__ norm) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    float s = 0.f;
    for (int32_t k = rowptr[w] + lane; k < rowptr[w + 1]; k += 32) s += dinv[colidx[k]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) norm[w] = s * dinv[w];
}
// val[k] = norm[row(k)] (forward: (A0 h) * norm) ; valT in CSR order for the aliased transpose: norm[colidx[k]]
__global__ void aswritten_val_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                     const float *__restrict__ norm, int32_t n_rows, float *__restrict__ val,
                                     float *__restrict__ valT_alias) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const float nr = norm[w];
    for (int32_t k = rowptr[w] + lane; k < rowptr[w + 1]; k += 32) {
        val[k] = nr;
        if (valT_alias) valT_alias[k] = norm[colidx[k]];
    }
}

__global__ void gather_f32_kernel(const float *__restrict__ src, const int32_t *__restrict__ idx, int64_t n,
                                  float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[idx[i]];
}

__global__ void to_dense_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                const float *__restrict__ val, int32_t n_rows, float *__restrict__ out, int64_t ld) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    for (int32_t k = rowptr[w] + lane; k < rowptr[w + 1]; k += 32) out[w * ld + colidx[k]] = val ? val[k] : 1.0f;
}

// flags[i] = int(A[i]) != 0 over the row-major dense matrix; flags[n] = 0 (scan sentinel)
__global__ void dense_flags_kernel(const float *__restrict__ A, int64_t rows, int64_t cols, int64_t ld,
                                   uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n = rows * cols;
    if (i > n) return;
    if (i == n) { flags[i] = 0; return; }
    const float a = A[(i / cols) * ld + (i % cols)];
    flags[i] = ((int)a != 0) ? 1u : 0u;
}
__global__ void dense_compact_kernel(const float *__restrict__ A, int64_t rows, int64_t cols, int64_t ld,
                                     const uint32_t *__restrict__ pos, int32_t *__restrict__ out_rows,
                                     int32_t *__restrict__ out_cols, float *__restrict__ out_vals) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const uint32_t p = pos[i];
    if (pos[i + 1] != p) {
        const int64_t r = i / cols, c = i % cols;
        out_rows[p] = (int32_t)r;
        out_cols[p] = (int32_t)c;
        if (out_vals) out_vals[p] = A[r * ld + c];
    }
}

__global__ void rebase_kernel(const int32_t *__restrict__ in, int64_t n, int32_t base, int32_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] - base;
}

static inline unsigned grid_for(int64_t n, int threads) { return (unsigned)ceil_div(n > 0 ? n : 1, threads); }

static int finish_stats(gnn_ctx *ctx, gnn_graph *g) {
    int32_t *d_max = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&d_max, 16, ctx->stream));
    const int32_t init[4] = {0, 0x7fffffff, 0, 0x7fffffff};
    GNN_CHECK_CUDA(cudaMemcpyAsync(d_max, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    max_diff_kernel<<<grid_for(g->n_rows, 256), 256, 0, ctx->stream>>>(g->rowptr, g->n_rows, d_max);
    GNN_LAUNCHED(ctx);
    if (g->colptr) {
        max_diff_kernel<<<grid_for(g->t_rows, 256), 256, 0, ctx->stream>>>(g->colptr, g->t_rows, d_max + 2);
        GNN_LAUNCHED(ctx);
    }
    int32_t h[4] = {0, 0, 0, 0};
    GNN_CHECK_CUDA(cudaMemcpyAsync(h, d_max, 16, cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    GNN_CHECK_CUDA(cudaFreeAsync(d_max, ctx->stream));
    g->max_row_nnz = h[0];
    g->min_row_nnz = g->n_rows > 0 ? h[1] : 0;
    if (g->colptr) {
        g->max_col_nnz = h[2];
        g->min_col_nnz = g->t_rows > 0 ? h[3] : 0;
    }
    return 0;
}

} // namespace gnn

using namespace gnn;

extern "C" {

int gnn_graph_build(gnn_ctx_t *ctx, const int32_t *src, const int32_t *dst, int64_t E, int32_t N, int fill_mode,
                    gnn_graph_t **out) {
    GNN_REQUIRE(ctx && out, "gnn_graph_build: NULL argument");
    GNN_REQUIRE(N > 0, "dims cannot be empty or zero");
    GNN_REQUIRE(E >= 0 && fill_mode >= 0 && fill_mode <= 2, "gnn_graph_build: bad E or fill_mode");
    GNN_REQUIRE(E == 0 || (src && dst), "gnn_graph_build: NULL edge arrays");
    const int cb = bits_for(N);
    const int64_t m = E + (fill_mode == 1 ? N : 0);
    GNN_REQUIRE(m < (int64_t)0x7FFFFFFF, "gnn_graph_build: E + N = %lld exceeds int32 positions", (long long)m);
    cudaStream_t s = ctx->stream;
    gnn_graph *g = new gnn_graph();
    g->n_rows = g->n_cols = N;
    g->t_rows = N;
    g->fill_mode = fill_mode;
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->rowptr, (size_t)(N + 1) * 4));
    if (m == 0) {
        GNN_CHECK_CUDA(cudaMemsetAsync(g->rowptr, 0, (size_t)(N + 1) * 4, s));
        GNN_CHECK_CUDA(cudaMalloc((void **)&g->colidx, 4));
        g->nnz = 0;
        *out = g;
        return 0;
    }
    uint64_t *keys = nullptr, *ukeys = nullptr;
    uint32_t *flags = nullptr;
    int *err_d = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&keys, (size_t)m * 8, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&flags, (size_t)(m + 2) * 4, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&err_d, 4, s));
    GNN_CHECK_CUDA(cudaMemsetAsync(err_d, 0, 4, s));
    make_keys_kernel<<<grid_for(m, 256), 256, 0, s>>>(src, dst, E, N, N, cb, fill_mode == 1, keys, err_d);
    GNN_LAUNCHED(ctx);
    GNN_TRY(radix_sort_u64(ctx, keys, nullptr, m, 0, 2 * cb));
    unique_flags_kernel<<<grid_for(m, 256), 256, 0, s>>>(keys, m, cb, fill_mode == 0, flags);
    GNN_LAUNCHED(ctx);
    // scan m+1 entries (flags[m] = 0) so that pos[m] = nnz and "kept" == pos[i+1] != pos[i] for every i
    GNN_CHECK_CUDA(cudaMemsetAsync(flags + m, 0, 8, s));
    GNN_TRY(exclusive_scan_u32(ctx, flags, flags, m + 1, nullptr));
    uint32_t h_nnz = 0;
    int h_err = 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(&h_nnz, flags + m, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaMemcpyAsync(&h_err, err_d, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    if (h_err) {
        cudaFreeAsync(keys, s); cudaFreeAsync(flags, s); cudaFreeAsync(err_d, s);
        cudaFree(g->rowptr);
        delete g;
        // same condition and message as graph::Data's constructor (reference src/graph.cpp:87-88)
        set_error("invalid input, max value in edge_index should be less than the number of nodes from x");
        return 2;
    }
    g->nnz = h_nnz;
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->colidx, (size_t)(g->nnz ? g->nnz : 1) * 4));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&ukeys, (size_t)(g->nnz ? g->nnz : 1) * 8, s));
    compact_kernel<<<grid_for(m, 256), 256, 0, s>>>(keys, flags, m, cb, ukeys, g->colidx);
    GNN_LAUNCHED(ctx);
    lower_bound_kernel<<<grid_for(N + 1, 256), 256, 0, s>>>(ukeys, g->nnz, cb, ((uint64_t)1 << (64 - cb)) - 1, N,
                                                           g->rowptr);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaFreeAsync(keys, s));
    GNN_CHECK_CUDA(cudaFreeAsync(ukeys, s));
    GNN_CHECK_CUDA(cudaFreeAsync(flags, s));
    GNN_CHECK_CUDA(cudaFreeAsync(err_d, s));
    GNN_TRY(finish_stats(ctx, g));
    *out = g;
    return 0;
}

int gnn_graph_build_h(gnn_ctx_t *ctx, const int32_t *src_h, const int32_t *dst_h, int64_t E, int32_t N, int fill_mode,
                      gnn_graph_t **out) {
    int32_t *d = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&d, (size_t)(E ? E : 1) * 8, ctx->stream));
    if (E) {
        GNN_CHECK_CUDA(cudaMemcpyAsync(d, src_h, (size_t)E * 4, cudaMemcpyHostToDevice, ctx->stream));
        GNN_CHECK_CUDA(cudaMemcpyAsync(d + E, dst_h, (size_t)E * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    int r = gnn_graph_build(ctx, d, d + E, E, N, fill_mode, out);
    cudaFreeAsync(d, ctx->stream);
    return r;
}

int gnn_graph_build_weighted(gnn_ctx_t *ctx, const int32_t *src, const int32_t *dst, const float *w, int64_t E, int32_t N,
                             int fill_mode, gnn_graph_t **out) {
    GNN_REQUIRE(ctx && out && src && dst && w, "gnn_graph_build_weighted: NULL argument");
    GNN_REQUIRE(N > 0, "dims cannot be empty or zero");
    GNN_REQUIRE(E > 0 && (fill_mode == 1 || fill_mode == 2), "gnn_graph_build_weighted: bad E or fill_mode (1 or 2)");
    const int cb = bits_for(N);
    const int64_t m = E + (fill_mode == 1 ? N : 0);
    GNN_REQUIRE(m < (int64_t)0x7FFFFFFF, "gnn_graph_build_weighted: E + N = %lld exceeds int32 positions", (long long)m);
    cudaStream_t s = ctx->stream;
    uint64_t *keys = nullptr, *ukeys = nullptr;
    uint32_t *pos_in = nullptr, *flags = nullptr;
    int *err_d = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&keys, (size_t)m * 8, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&pos_in, (size_t)m * 4, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&flags, (size_t)(m + 2) * 4, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&err_d, 4, s));
    GNN_CHECK_CUDA(cudaMemsetAsync(err_d, 0, 4, s));
    make_keys_w_kernel<<<grid_for(m, 256), 256, 0, s>>>(src, dst, E, N, cb, fill_mode == 1, keys, pos_in, err_d);
    GNN_LAUNCHED(ctx);
    GNN_TRY(radix_sort_u64(ctx, keys, pos_in, m, 0, 2 * cb)); // stable: equal keys keep edge order
    last_flags_kernel<<<grid_for(m, 256), 256, 0, s>>>(keys, m, flags);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaMemsetAsync(flags + m, 0, 8, s));
    GNN_TRY(exclusive_scan_u32(ctx, flags, flags, m + 1, nullptr));
    uint32_t h_nnz = 0;
    int h_err = 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(&h_nnz, flags + m, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaMemcpyAsync(&h_err, err_d, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    if (h_err) {
        cudaFreeAsync(keys, s); cudaFreeAsync(pos_in, s); cudaFreeAsync(flags, s); cudaFreeAsync(err_d, s);
        set_error("invalid input, max value in edge_index should be less than the number of nodes from x");
        return 2;
    }
    gnn_graph *g = new gnn_graph();
    g->n_rows = g->n_cols = N;
    g->t_rows = N;
    g->fill_mode = fill_mode;
    g->nnz = h_nnz;
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->rowptr, (size_t)(N + 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->colidx, (size_t)(g->nnz ? g->nnz : 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->val0, (size_t)(g->nnz ? g->nnz : 1) * 4));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&ukeys, (size_t)(g->nnz ? g->nnz : 1) * 8, s));
    compact_w_kernel<<<grid_for(m, 256), 256, 0, s>>>(keys, pos_in, flags, m, cb, E, w, ukeys, g->colidx, g->val0);
    GNN_LAUNCHED(ctx);
    lower_bound_kernel<<<grid_for(N + 1, 256), 256, 0, s>>>(ukeys, g->nnz, cb, ((uint64_t)1 << (64 - cb)) - 1, N, g->rowptr);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaFreeAsync(keys, s));
    GNN_CHECK_CUDA(cudaFreeAsync(ukeys, s));
    GNN_CHECK_CUDA(cudaFreeAsync(pos_in, s));
    GNN_CHECK_CUDA(cudaFreeAsync(flags, s));
    GNN_CHECK_CUDA(cudaFreeAsync(err_d, s));
    GNN_TRY(finish_stats(ctx, g));
    *out = g;
    return 0;
}

int gnn_graph_export_weights_h(gnn_ctx_t *ctx, const gnn_graph_t *g, float *val0_h) {
    GNN_REQUIRE(ctx && g && val0_h && g->val0, "gnn_graph_export_weights_h: graph has no raw weights (gnn_graph_build_weighted)");
    GNN_CHECK_CUDA(cudaMemcpyAsync(val0_h, g->val0, (size_t)g->nnz * 4, cudaMemcpyDeviceToHost, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int gnn_graph_from_csr(gnn_ctx_t *ctx, int32_t n_rows, int32_t n_cols, const int32_t *rowptr, const int32_t *colidx,
                       const float *val, gnn_graph_t **out) {
    GNN_REQUIRE(ctx && out && rowptr, "gnn_graph_from_csr: NULL argument");
    GNN_REQUIRE(n_rows > 0 && n_cols > 0, "dims cannot be empty or zero");
    cudaStream_t s = ctx->stream;
    int32_t ends[2];
    GNN_CHECK_CUDA(cudaMemcpyAsync(&ends[0], rowptr, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaMemcpyAsync(&ends[1], rowptr + n_rows, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    gnn_graph *g = new gnn_graph();
    g->n_rows = n_rows; g->n_cols = n_cols; g->t_rows = n_cols; g->fill_mode = 2;
    g->nnz = ends[1] - ends[0];
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->rowptr, (size_t)(n_rows + 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->colidx, (size_t)(g->nnz ? g->nnz : 1) * 4));
    rebase_kernel<<<grid_for(n_rows + 1, 256), 256, 0, s>>>(rowptr, n_rows + 1, ends[0], g->rowptr);
    GNN_LAUNCHED(ctx);
    if (g->nnz) GNN_CHECK_CUDA(cudaMemcpyAsync(g->colidx, colidx + ends[0], (size_t)g->nnz * 4, cudaMemcpyDeviceToDevice, s));
    if (val) {
        GNN_CHECK_CUDA(cudaMalloc((void **)&g->val, (size_t)(g->nnz ? g->nnz : 1) * 4));
        if (g->nnz) GNN_CHECK_CUDA(cudaMemcpyAsync(g->val, val + ends[0], (size_t)g->nnz * 4, cudaMemcpyDeviceToDevice, s));
    }
    GNN_TRY(finish_stats(ctx, g));
    *out = g;
    return 0;
}

int gnn_graph_build_csc(gnn_ctx_t *ctx, gnn_graph_t *g) {
    GNN_REQUIRE(ctx && g, "gnn_graph_build_csc: NULL argument");
    if (g->colptr) return 0;
    cudaStream_t s = ctx->stream;
    const int cb = bits_for(g->n_cols);
    const int64_t nnz = g->nnz;
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->colptr, (size_t)(g->n_cols + 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->rowidx, (size_t)(nnz ? nnz : 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&g->perm, (size_t)(nnz ? nnz : 1) * 4));
    if (nnz == 0) {
        GNN_CHECK_CUDA(cudaMemsetAsync(g->colptr, 0, (size_t)(g->n_cols + 1) * 4, s));
        g->symmetric = g->n_rows == g->n_cols;
        return 0;
    }
    uint64_t *keys = nullptr;
    uint32_t *vals = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&keys, (size_t)nnz * 8, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&vals, (size_t)nnz * 4, s));
    csr_to_keys_kernel<<<grid_for((int64_t)g->n_rows * 32, 256), 256, 0, s>>>(g->rowptr, g->colidx, g->n_rows, cb, keys,
                                                                              vals);
    GNN_LAUNCHED(ctx);
    // CSR order is (row, col) ascending; a STABLE sort on the column bits alone yields (col, row) ascending
    GNN_TRY(radix_sort_u64(ctx, keys, vals, nnz, 0, cb));
    unpack_csc_kernel<<<grid_for(nnz, 256), 256, 0, s>>>(keys, vals, nnz, cb, g->rowidx, g->perm);
    GNN_LAUNCHED(ctx);
    lower_bound_kernel<<<grid_for(g->n_cols + 1, 256), 256, 0, s>>>(keys, nnz, 0, ((uint64_t)1 << cb) - 1, g->n_cols,
                                                                   g->colptr);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaFreeAsync(keys, s));
    GNN_CHECK_CUDA(cudaFreeAsync(vals, s));
    // structural symmetry: CSC arrays identical to CSR arrays -> alias them (halves index traffic in L2)
    if (g->n_rows == g->n_cols) {
        int *differ = nullptr, h = 0;
        GNN_CHECK_CUDA(cudaMallocAsync((void **)&differ, 4, s));
        GNN_CHECK_CUDA(cudaMemsetAsync(differ, 0, 4, s));
        arrays_equal_kernel<<<grid_for(g->n_rows + 1, 256), 256, 0, s>>>(g->rowptr, g->colptr, g->n_rows + 1, differ);
        GNN_LAUNCHED(ctx);
        arrays_equal_kernel<<<grid_for(nnz, 256), 256, 0, s>>>(g->colidx, g->rowidx, nnz, differ);
        GNN_LAUNCHED(ctx);
        GNN_CHECK_CUDA(cudaMemcpyAsync(&h, differ, 4, cudaMemcpyDeviceToHost, s));
        GNN_CHECK_CUDA(cudaStreamSynchronize(s));
        GNN_CHECK_CUDA(cudaFreeAsync(differ, s));
        g->symmetric = (h == 0);
    }
    g->nnz_t = g->nnz;
    GNN_TRY(finish_stats(ctx, g));
    return 0;
}

int gnn_graph_normalize(gnn_ctx_t *ctx, gnn_graph_t *g) {
    GNN_REQUIRE(ctx && g, "gnn_graph_normalize: NULL argument");
    GNN_REQUIRE(g->n_rows == g->n_cols, "gnn_graph_normalize: needs the square (global) graph");
    cudaStream_t s = ctx->stream;
    const int64_t nnz = g->nnz ? g->nnz : 1;
    if (!g->deg) GNN_CHECK_CUDA(cudaMalloc((void **)&g->deg, (size_t)g->n_rows * 4));
    if (!g->dinv) GNN_CHECK_CUDA(cudaMalloc((void **)&g->dinv, (size_t)g->n_rows * 4));
    if (!g->val) GNN_CHECK_CUDA(cudaMalloc((void **)&g->val, (size_t)nnz * 4));
    if (g->val0) { // weighted adjacency: weighted degree, val = (w * dinv[r]) * dinv[c]
        degree_w_kernel<<<grid_for((int64_t)g->n_rows * 32, 256), 256, 0, s>>>(g->rowptr, g->val0, g->n_rows, g->deg, g->dinv);
        GNN_LAUNCHED(ctx);
        edge_val_w_kernel<<<grid_for((int64_t)g->n_rows * 32, 256), 256, 0, s>>>(g->rowptr, g->colidx, g->val0, g->dinv,
                                                                                g->n_rows, g->val);
        GNN_LAUNCHED(ctx);
    } else {
        degree_kernel<<<grid_for(g->n_rows, 256), 256, 0, s>>>(g->rowptr, g->n_rows, g->deg, g->dinv);
        GNN_LAUNCHED(ctx);
        edge_val_kernel<<<grid_for((int64_t)g->n_rows * 32, 256), 256, 0, s>>>(g->rowptr, g->colidx, g->dinv, g->n_rows,
                                                                               g->val);
        GNN_LAUNCHED(ctx);
    }
    if (g->colptr && !g->valT) {
        GNN_CHECK_CUDA(cudaMalloc((void **)&g->valT, (size_t)nnz * 4));
        gather_f32_kernel<<<grid_for(g->nnz, 256), 256, 0, s>>>(g->val, g->perm, g->nnz, g->valT);
        GNN_LAUNCHED(ctx);
    }
    return 0;
}

int gnn_graph_normalize_as_written(gnn_ctx_t *ctx, gnn_graph_t *g, float *norm_out) {
    GNN_REQUIRE(ctx && g, "gnn_graph_normalize_as_written: NULL argument");
    GNN_REQUIRE(g->n_rows == g->n_cols, "gnn_graph_normalize_as_written: needs the square (global) graph");
    GNN_REQUIRE(g->fill_mode == 0, "gnn_graph_normalize_as_written: build the graph with fill_mode 0 (add_self_loops(.., 0))");
    cudaStream_t s = ctx->stream;
    const int64_t nnz = g->nnz ? g->nnz : 1;
    if (!g->deg) GNN_CHECK_CUDA(cudaMalloc((void **)&g->deg, (size_t)g->n_rows * 4));
    if (!g->dinv) GNN_CHECK_CUDA(cudaMalloc((void **)&g->dinv, (size_t)g->n_rows * 4));
    if (!g->val) GNN_CHECK_CUDA(cudaMalloc((void **)&g->val, (size_t)nnz * 4));
    if (g->colptr && !g->valT) GNN_CHECK_CUDA(cudaMalloc((void **)&g->valT, (size_t)nnz * 4));
    float *norm = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&norm, (size_t)g->n_rows * 4, s));
    aswritten_dinv_kernel<<<grid_for(g->n_rows, 256), 256, 0, s>>>(g->rowptr, g->n_rows, g->deg, g->dinv);
    GNN_LAUNCHED(ctx);
    aswritten_norm_kernel<<<grid_for((int64_t)g->n_rows * 32, 256), 256, 0, s>>>(g->rowptr, g->colidx, g->dinv, g->n_rows, norm);
    GNN_LAUNCHED(ctx);
    // the values are constant along a row, so the transpose differs from the matrix even when the structure is
    // symmetric: the aliased (CSR-as-CSC) backward gets its own value array in CSR order
    aswritten_val_kernel<<<grid_for((int64_t)g->n_rows * 32, 256), 256, 0, s>>>(g->rowptr, g->colidx, norm, g->n_rows, g->val,
                                                                            (g->colptr && g->symmetric) ? g->valT : nullptr);
    GNN_LAUNCHED(ctx);
    if (g->colptr && !g->symmetric) {
        gather_f32_kernel<<<grid_for(g->nnz, 256), 256, 0, s>>>(g->val, g->perm, g->nnz, g->valT);
        GNN_LAUNCHED(ctx);
    }
    if (norm_out) GNN_CHECK_CUDA(cudaMemcpyAsync(norm_out, norm, (size_t)g->n_rows * 4, cudaMemcpyDeviceToDevice, s));
    GNN_CHECK_CUDA(cudaFreeAsync(norm, s));
    return 0;
}

int gnn_graph_destroy(gnn_ctx_t *ctx, gnn_graph_t *g) {
    if (!g) return 0;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    // (graph arrays are plain cudaMalloc allocations: taking them from the stream-ordered pool was measured and made a
    // rebuilt products-shaped structure 4x slower — 130 ms instead of 30 — because the pool grows by mapping fresh
    // physical memory while the previous structure is still alive; only the sort temporaries live in the pool)
    cudaFree(g->rowptr); cudaFree(g->colidx); cudaFree(g->val); cudaFree(g->val0);
    cudaFree(g->colptr); cudaFree(g->rowidx); cudaFree(g->perm); cudaFree(g->valT);
    cudaFree(g->deg); cudaFree(g->dinv);
    delete g;
    return 0;
}

int64_t gnn_graph_nnz(const gnn_graph_t *g) { return g ? g->nnz : -1; }
int32_t gnn_graph_rows(const gnn_graph_t *g) { return g ? g->n_rows : -1; }
int32_t gnn_graph_cols(const gnn_graph_t *g) { return g ? g->n_cols : -1; }
int gnn_graph_is_symmetric(const gnn_graph_t *g) { return g && g->symmetric; }

int gnn_graph_export_h(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t *rowptr_h, int32_t *colidx_h, float *val_h,
                       int32_t *colptr_h, int32_t *rowidx_h, int32_t *perm_h, float *valT_h, int32_t *deg_h,
                       float *dinv_h) {
    GNN_REQUIRE(ctx && g, "gnn_graph_export_h: NULL argument");
    cudaStream_t s = ctx->stream;
    const size_t nz = (size_t)g->nnz;
#define EXPORT(dst, srcp, bytes, what)                                                          \
    if (dst) {                                                                                  \
        GNN_REQUIRE((srcp) != nullptr, "gnn_graph_export_h: %s not built", what);               \
        if (bytes) GNN_CHECK_CUDA(cudaMemcpyAsync(dst, srcp, bytes, cudaMemcpyDeviceToHost, s)); \
    }
    EXPORT(rowptr_h, g->rowptr, (size_t)(g->n_rows + 1) * 4, "rowptr");
    EXPORT(colidx_h, g->colidx, nz * 4, "colidx");
    EXPORT(val_h, g->val, nz * 4, "val");
    EXPORT(colptr_h, g->colptr, (size_t)(g->t_rows + 1) * 4, "colptr (call gnn_graph_build_csc)");
    EXPORT(rowidx_h, g->rowidx, nz * 4, "rowidx (call gnn_graph_build_csc)");
    EXPORT(perm_h, g->perm, nz * 4, "perm (call gnn_graph_build_csc)");
    EXPORT(valT_h, g->valT, nz * 4, "valT (call gnn_graph_build_csc then gnn_graph_normalize)");
    EXPORT(deg_h, g->deg, (size_t)g->n_rows * 4, "deg (call gnn_graph_normalize)");
    EXPORT(dinv_h, g->dinv, (size_t)g->n_rows * 4, "dinv (call gnn_graph_normalize)");
#undef EXPORT
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int gnn_graph_to_dense(gnn_ctx_t *ctx, const gnn_graph_t *g, int weighted, float *out, int64_t ld) {
    GNN_REQUIRE(ctx && g && out, "gnn_graph_to_dense: NULL argument");
    GNN_REQUIRE(weighted != 1 || g->val, "gnn_graph_to_dense: values not built (call gnn_graph_normalize)");
    GNN_REQUIRE(weighted != 2 || g->val0, "gnn_graph_to_dense: graph has no raw weights (gnn_graph_build_weighted)");
    GNN_CHECK_CUDA(cudaMemset2DAsync(out, (size_t)ld * 4, 0, (size_t)g->n_cols * 4, g->n_rows, ctx->stream));
    to_dense_kernel<<<grid_for((int64_t)g->n_rows * 32, 256), 256, 0, ctx->stream>>>(g->rowptr, g->colidx,
                                                                                    weighted == 2 ? g->val0 : (weighted ? g->val : nullptr),
                                                                                    g->n_rows, out, ld);
    GNN_LAUNCHED(ctx);
    return 0;
}

// dense -> row-major COO of the entries with int(a) != 0  (graph::adj_to_edge_list, reference src/graph.cpp:46-67)
int gnn_dense_to_coo(gnn_ctx_t *ctx, const float *A, int64_t rows, int64_t cols, int64_t ld, int32_t *out_rows,
                     int32_t *out_cols, float *out_vals, int64_t capacity, int64_t *count_h) {
    GNN_REQUIRE(ctx && A && count_h && rows > 0 && cols > 0, "gnn_dense_to_coo: bad argument");
    const int64_t n = rows * cols;
    GNN_REQUIRE(n < (int64_t)0x7FFFFFFF, "gnn_dense_to_coo: matrix too large for the dense path");
    cudaStream_t s = ctx->stream;
    uint32_t *flags = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&flags, (size_t)(n + 1) * 4, s));
    dense_flags_kernel<<<grid_for(n + 1, 256), 256, 0, s>>>(A, rows, cols, ld, flags);
    GNN_LAUNCHED(ctx);
    GNN_TRY(exclusive_scan_u32(ctx, flags, flags, n + 1, nullptr));
    uint32_t cnt = 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(&cnt, flags + n, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    *count_h = cnt;
    if (out_rows && out_cols && cnt > 0) {
        GNN_REQUIRE((int64_t)cnt <= capacity, "gnn_dense_to_coo: %u entries exceed the output capacity %lld", cnt,
                    (long long)capacity);
        dense_compact_kernel<<<grid_for(n, 256), 256, 0, s>>>(A, rows, cols, ld, flags, out_rows, out_cols, out_vals);
        GNN_LAUNCHED(ctx);
    }
    GNN_CHECK_CUDA(cudaFreeAsync(flags, s));
    return 0;
}

int gnn_partition_ptr_h(int64_t N, int32_t P, int64_t *part_ptr_h) {
    GNN_REQUIRE(N > 0 && P > 0 && part_ptr_h, "gnn_partition_ptr_h: bad argument");
    const int64_t chunk = (N + P - 1) / P;
    for (int32_t p = 0; p <= P; p++) {
        const int64_t v = (int64_t)p * chunk;
        part_ptr_h[p] = v < N ? v : N;
    }
    return 0;
}

int gnn_graph_slice_rows(gnn_ctx_t *ctx, const gnn_graph_t *g, int64_t lo, int64_t hi, gnn_graph_t **out) {
    GNN_REQUIRE(ctx && g && out, "gnn_graph_slice_rows: NULL argument");
    GNN_REQUIRE(0 <= lo && lo < hi && hi <= g->n_rows, "gnn_graph_slice_rows: bad range [%lld,%lld)", (long long)lo,
                (long long)hi);
    GNN_REQUIRE(g->n_rows == g->n_cols, "gnn_graph_slice_rows: needs the square (global) graph");
    gnn_graph *l = nullptr;
    // forward block: rows [lo,hi) of A_hat (global columns)
    GNN_TRY(gnn_graph_from_csr(ctx, (int32_t)(hi - lo), g->n_cols, g->rowptr + lo, g->colidx, g->val, &l));
    // backward block: rows [lo,hi) of A_hat^T = CSC columns [lo,hi) (global row ids)
    if (g->colptr) {
        gnn_graph *t = nullptr;
        const int32_t *tptr = g->symmetric ? g->rowptr : g->colptr;
        const int32_t *tidx = g->symmetric ? g->colidx : g->rowidx;
        // the as-written normalisation keeps a transposed value array even on a symmetric structure (its values are
        // constant per row, so not symmetric): same rule as spmm_rows_range / gnn_spmm_bwd
        const float *tval = g->symmetric ? (g->valT ? g->valT : g->val) : g->valT;
        GNN_TRY(gnn_graph_from_csr(ctx, (int32_t)(hi - lo), g->n_rows, tptr + lo, tidx, tval, &t));
        l->colptr = t->rowptr; l->rowidx = t->colidx; l->valT = t->val;
        l->max_col_nnz = t->max_row_nnz;
        l->min_col_nnz = t->min_row_nnz;
        l->t_rows = t->n_rows;
        l->nnz_t = t->nnz;
        t->rowptr = nullptr; t->colidx = nullptr; t->val = nullptr;
        delete t;
    }
    *out = l;
    return 0;
}

} // extern "C"
