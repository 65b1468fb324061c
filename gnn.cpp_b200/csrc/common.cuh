// common.cuh — shared host-side plumbing for libgnn_b200.so (context, error reporting, launch
// accounting, workspace).  sm_100a only; there is no CPU fallback anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/gnn_c.h"

namespace gnn {

void set_error(const char *fmt, ...);

#define GNN_CHECK_CUDA(expr)                                                                          \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            gnn::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));     \
            return 1;                                                                                 \
        }                                                                                             \
    } while (0)

#define GNN_REQUIRE(cond, ...)                                                                        \
    do {                                                                                              \
        if (!(cond)) {                                                                                \
            gnn::set_error(__VA_ARGS__);                                                              \
            return 2;                                                                                 \
        }                                                                                             \
    } while (0)

#define GNN_TRY(expr)                                                                                 \
    do {                                                                                              \
        int _r = (expr);                                                                              \
        if (_r) return _r;                                                                            \
    } while (0)

// after a kernel launch: count it and surface launch-configuration errors
#define GNN_LAUNCHED(ctx)                                                                             \
    do {                                                                                              \
        (ctx)->launches++;                                                                            \
        GNN_CHECK_CUDA(cudaGetLastError());                                                           \
    } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

} // namespace gnn

struct gnn_ctx {
    int device = 0;
    int sm_count = 148;
    size_t l2_bytes = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    int spmm_variant = 0, spmm_tune_u = 0, spmm_tune_pf = 1, spmm_chunk = 0;
    int spmm_alt = 0; // lane mapping for non-power-of-two vector counts (GNN_SPMM_ALT, see spmm_launch)
    int spmm_async = 0; // FIFO depth of the shared-memory-staged merge kernel for widths <= 128 (GNN_SPMM_ASYNC; 0 = register gathers)
    // grow-only scratch (sort double buffers, scan levels, split-K partials ...)
    void *ws = nullptr;
    size_t ws_bytes = 0;
    uint64_t ws_gen = 0; // bumped whenever `ws` is reallocated: captured CUDA graphs that bake ws pointers check it
    // NCCL (comm.cu)
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;

    // returns a scratch region of at least `bytes`; contents are undefined. Stream-ordered growth.
    int workspace(size_t bytes, void **out);
};

struct gnn_graph {
    int32_t n_rows = 0, n_cols = 0;
    int64_t nnz = 0;
    // transposed block (CSC): t_rows rows of A^T (= n_cols for a whole graph, = n_rows for a row slice)
    int32_t t_rows = 0;
    int64_t nnz_t = 0;
    int fill_mode = 1;
    // CSR
    int32_t *rowptr = nullptr, *colidx = nullptr;
    float *val = nullptr;
    float *val0 = nullptr; // raw edge weights of a weighted adjacency (gnn_graph_build_weighted), NULL for 0/1 graphs
    // CSC (CSR of the transpose); aliases the CSR arrays when symmetric
    int32_t *colptr = nullptr, *rowidx = nullptr, *perm = nullptr;
    float *valT = nullptr;
    bool symmetric = false;
    // normalisation
    int32_t *deg = nullptr;
    float *dinv = nullptr;
    // degree statistics (host) for SpMM variant selection
    int32_t max_row_nnz = 0, max_col_nnz = 0, min_row_nnz = 0, min_col_nnz = 0;
};

namespace gnn {
// sort_scan.cu ------------------------------------------------------------------------------------
// exclusive scan of n uint32 values (in -> out, may alias); total (optional, device) = sum of all.
int exclusive_scan_u32(gnn_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total_d);
// stable LSD radix sort of 64-bit keys on bits [bit_lo, bit_hi); optional 32-bit payload.
// keys/vals are overwritten with the sorted sequence (scratch comes from ctx->workspace).
int radix_sort_u64(gnn_ctx *ctx, uint64_t *keys, uint32_t *vals, int64_t n, int bit_lo, int bit_hi);
// spmm.cu -----------------------------------------------------------------------------------------
// Where the output rows of an aggregation go.  rows_per == 0: the plain matrix Y (base[0] = Y).  Otherwise output row
// r belongs to destination q = r / rows_per and is written to base[q] + (r - q * rows_per) * ldy: the 2-D partitioned
// trainer points base[q] at peer q's IPC-mapped activation buffer, so the aggregation kernel itself performs the
// "columns -> rows" exchange with its epilogue stores (compute and transfer fused in one kernel).
constexpr int SPMM_MAX_DEST = 8;
struct YDest {
    float *base[SPMM_MAX_DEST];
    int32_t rows_per;
    const int32_t *row_map; // optional: kernel row i is output row row_map[i] (aggregation over a row SUBSET: the interior /
                            // boundary split of the partitioned trainer); nullptr = identity
};
int spmm_launch(gnn_ctx *ctx, int32_t n_out, int64_t k_base, int64_t nnz, const int32_t *ptr, const int32_t *idx,
                const float *val, int32_t min_nnz_row, int32_t max_nnz_row, const float *P, int64_t ldp, int32_t F,
                float *Y, int64_t ldy, const float *bias, int relu, const float *mask, int64_t ldm,
                bool may_touch_padding = true, const YDest *dest = nullptr);
int spmm_rows_range(gnn_ctx *ctx, const gnn_graph *g, int transpose, int32_t r0, int32_t r1, int64_t k0, int64_t k1,
                    const float *P, int64_t ldp, int32_t F, float *Y, int64_t ldy, const float *bias, int relu,
                    const float *mask, int64_t ldm, const YDest *dest = nullptr);
} // namespace gnn
