// partition.cu — per-rank arrays of the 1-D row partition, built on the device and bit-exact against
// the CPU restatement (oracle/gcn_oracle.c: orc_partition_halo, orc_partition_interior):
//   halo_ids        sorted unique column ids of the rank's rows that lie outside its own row range [lo, hi)
//   local_colidx    the rank's column array renumbered for a [own rows ; halo rows] input matrix: an owned column c
//                   becomes c - lo, a halo column n_loc + (position of c in halo_ids)
//   interior flag   1 for a row whose columns are all owned (it can be aggregated before any halo row arrives),
//                   interior_rows / boundary_rows: the two row lists, ascending
// The reference has no partitioning at all (it has no sparse or multi-device code); the aggregation being sharded is
// src/graph.cpp:204-212.
#include "common.cuh"

struct gnn_partition {
    int64_t lo = 0, hi = 0, nnz = 0;
    int32_t n_rows = 0, n_cols = 0;
    int64_t n_halo = 0, n_interior = 0;
    int32_t *halo_ids = nullptr, *local_colidx = nullptr, *interior_rows = nullptr, *boundary_rows = nullptr;
    uint8_t *interior = nullptr;
};

namespace gnn {

static inline unsigned pgrid(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, 256); }

__global__ void halo_mark_kernel(const int32_t *__restrict__ colidx, int64_t k0, int64_t nnz, int32_t lo, int32_t hi,
                                 uint32_t *__restrict__ flag) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int32_t c = colidx[k0 + k];
    if (c < lo || c >= hi) flag[c] = 1u; // benign race: every writer stores 1
}
__global__ void halo_compact_kernel(const uint32_t *__restrict__ flag, const uint32_t *__restrict__ pos, int32_t n_cols,
                                    int32_t *__restrict__ halo_ids) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_cols && flag[c]) halo_ids[pos[c]] = (int32_t)c;
}
__global__ void renumber_kernel(const int32_t *__restrict__ colidx, int64_t k0, int64_t nnz, int32_t lo, int32_t hi,
                                const uint32_t *__restrict__ pos, int32_t *__restrict__ local_colidx) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    const int32_t c = colidx[k0 + k];
    local_colidx[k] = (c >= lo && c < hi) ? c - lo : (hi - lo) + (int32_t)pos[c];
}
// one warp per row: interior iff no column leaves [lo, hi)
__global__ void interior_kernel(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, int32_t n_rows,
                                int32_t lo, int32_t hi, uint8_t *__restrict__ interior, uint32_t *__restrict__ iflag,
                                uint32_t *__restrict__ bflag) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    bool in = true;
    for (int32_t k = rowptr[row] + lane; k < rowptr[row + 1]; k += 32) {
        const int32_t c = colidx[k];
        if (c < lo || c >= hi) in = false;
    }
    in = __all_sync(0xffffffffu, in);
    if (lane == 0) {
        interior[row] = in ? 1 : 0;
        iflag[row] = in ? 1u : 0u;
        bflag[row] = in ? 0u : 1u;
    }
}
__global__ void rows_compact_kernel(const uint8_t *__restrict__ interior, const uint32_t *__restrict__ ipos,
                                    const uint32_t *__restrict__ bpos, int32_t n_rows, int32_t *__restrict__ interior_rows,
                                    int32_t *__restrict__ boundary_rows) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    if (interior[r]) interior_rows[ipos[r]] = (int32_t)r;
    else boundary_rows[bpos[r]] = (int32_t)r;
}

} // namespace gnn

using namespace gnn;

extern "C" {

int gnn_partition_build(gnn_ctx_t *ctx, const gnn_graph_t *g, int64_t lo, int64_t hi, int transpose, gnn_partition_t **out) {
    GNN_REQUIRE(ctx && g && out, "gnn_partition_build: NULL argument");
    const int32_t *ptr = transpose ? g->colptr : g->rowptr, *idx = transpose ? g->rowidx : g->colidx;
    const int32_t n_rows = transpose ? g->t_rows : g->n_rows;
    GNN_REQUIRE(ptr && idx, "gnn_partition_build: the %s block is not built", transpose ? "backward (CSC)" : "forward (CSR)");
    GNN_REQUIRE(0 <= lo && lo <= hi && hi - lo == n_rows && hi <= g->n_cols,
                "gnn_partition_build: rows [%lld,%lld) do not match the block (%d rows of %d nodes)", (long long)lo,
                (long long)hi, n_rows, g->n_cols);
    cudaStream_t s = ctx->stream;
    gnn_partition *p = new gnn_partition();
    p->lo = lo; p->hi = hi; p->n_rows = n_rows; p->n_cols = g->n_cols;
    int32_t ends[2] = {0, 0}; // a slice keeps absolute offsets rebased to 0; read them back to be safe
    GNN_CHECK_CUDA(cudaMemcpyAsync(&ends[0], ptr, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaMemcpyAsync(&ends[1], ptr + n_rows, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    const int64_t k0 = ends[0];
    p->nnz = ends[1] - ends[0];
    const int32_t N = g->n_cols;
    uint32_t *flag = nullptr, *pos = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&flag, (size_t)(N + 1) * 4, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&pos, (size_t)(N + 1) * 4, s));
    GNN_CHECK_CUDA(cudaMemsetAsync(flag, 0, (size_t)(N + 1) * 4, s));
    halo_mark_kernel<<<pgrid(p->nnz), 256, 0, s>>>(idx, k0, p->nnz, (int32_t)lo, (int32_t)hi, flag);
    GNN_LAUNCHED(ctx);
    GNN_TRY(exclusive_scan_u32(ctx, flag, pos, N + 1, nullptr));
    uint32_t n_halo = 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(&n_halo, pos + N, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    p->n_halo = n_halo;
    GNN_CHECK_CUDA(cudaMalloc((void **)&p->halo_ids, (size_t)(n_halo ? n_halo : 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&p->local_colidx, (size_t)(p->nnz ? p->nnz : 1) * 4));
    halo_compact_kernel<<<pgrid(N), 256, 0, s>>>(flag, pos, N, p->halo_ids);
    GNN_LAUNCHED(ctx);
    renumber_kernel<<<pgrid(p->nnz), 256, 0, s>>>(idx, k0, p->nnz, (int32_t)lo, (int32_t)hi, pos, p->local_colidx);
    GNN_LAUNCHED(ctx);
    // interior / boundary rows
    uint32_t *iflag = nullptr, *bflag = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&iflag, (size_t)(n_rows + 1) * 4, s));
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&bflag, (size_t)(n_rows + 1) * 4, s));
    GNN_CHECK_CUDA(cudaMemsetAsync(iflag, 0, (size_t)(n_rows + 1) * 4, s));
    GNN_CHECK_CUDA(cudaMemsetAsync(bflag, 0, (size_t)(n_rows + 1) * 4, s));
    GNN_CHECK_CUDA(cudaMalloc((void **)&p->interior, (size_t)(n_rows ? n_rows : 1)));
    // the slice's ptr array keeps the parent's absolute offsets only when it aliases it; gnn_graph_slice_rows rebases,
    // so offsets index idx directly either way
    interior_kernel<<<pgrid((int64_t)n_rows * 32), 256, 0, s>>>(ptr, idx, n_rows, (int32_t)lo, (int32_t)hi, p->interior, iflag, bflag);
    GNN_LAUNCHED(ctx);
    GNN_TRY(exclusive_scan_u32(ctx, iflag, iflag, n_rows + 1, nullptr));
    GNN_TRY(exclusive_scan_u32(ctx, bflag, bflag, n_rows + 1, nullptr));
    uint32_t n_int = 0;
    GNN_CHECK_CUDA(cudaMemcpyAsync(&n_int, iflag + n_rows, 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    p->n_interior = n_int;
    GNN_CHECK_CUDA(cudaMalloc((void **)&p->interior_rows, (size_t)(n_int ? n_int : 1) * 4));
    GNN_CHECK_CUDA(cudaMalloc((void **)&p->boundary_rows, (size_t)(n_rows - n_int ? n_rows - n_int : 1) * 4));
    rows_compact_kernel<<<pgrid(n_rows), 256, 0, s>>>(p->interior, iflag, bflag, n_rows, p->interior_rows, p->boundary_rows);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaFreeAsync(flag, s));
    GNN_CHECK_CUDA(cudaFreeAsync(pos, s));
    GNN_CHECK_CUDA(cudaFreeAsync(iflag, s));
    GNN_CHECK_CUDA(cudaFreeAsync(bflag, s));
    *out = p;
    return 0;
}

int64_t gnn_partition_halo_count(const gnn_partition_t *p) { return p ? p->n_halo : -1; }
int64_t gnn_partition_interior_count(const gnn_partition_t *p) { return p ? p->n_interior : -1; }
int64_t gnn_partition_nnz(const gnn_partition_t *p) { return p ? p->nnz : -1; }

int gnn_partition_export_h(gnn_ctx_t *ctx, const gnn_partition_t *p, int32_t *halo_ids_h, int32_t *local_colidx_h,
                           uint8_t *interior_h, int32_t *interior_rows_h, int32_t *boundary_rows_h) {
    GNN_REQUIRE(ctx && p, "gnn_partition_export_h: NULL argument");
    cudaStream_t s = ctx->stream;
    if (halo_ids_h && p->n_halo) GNN_CHECK_CUDA(cudaMemcpyAsync(halo_ids_h, p->halo_ids, (size_t)p->n_halo * 4, cudaMemcpyDeviceToHost, s));
    if (local_colidx_h && p->nnz) GNN_CHECK_CUDA(cudaMemcpyAsync(local_colidx_h, p->local_colidx, (size_t)p->nnz * 4, cudaMemcpyDeviceToHost, s));
    if (interior_h && p->n_rows) GNN_CHECK_CUDA(cudaMemcpyAsync(interior_h, p->interior, (size_t)p->n_rows, cudaMemcpyDeviceToHost, s));
    if (interior_rows_h && p->n_interior)
        GNN_CHECK_CUDA(cudaMemcpyAsync(interior_rows_h, p->interior_rows, (size_t)p->n_interior * 4, cudaMemcpyDeviceToHost, s));
    if (boundary_rows_h && p->n_rows - p->n_interior)
        GNN_CHECK_CUDA(cudaMemcpyAsync(boundary_rows_h, p->boundary_rows, (size_t)(p->n_rows - p->n_interior) * 4, cudaMemcpyDeviceToHost, s));
    GNN_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}

int gnn_partition_destroy(gnn_ctx_t *ctx, gnn_partition_t *p) {
    if (!p) return 0;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    cudaFree(p->halo_ids); cudaFree(p->local_colidx); cudaFree(p->interior); cudaFree(p->interior_rows); cudaFree(p->boundary_rows);
    delete p;
    return 0;
}

} // extern "C"
