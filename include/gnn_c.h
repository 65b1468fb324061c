/* gnn_c.h — C ABI of the B200-native GCN hot path (libgnn_b200.so).
 *
 * This is the drop-in boundary: the reference (walexi/gnn.cpp) has no FFI of its own — its intended but
 * unwritten seam is a `device::` namespace called from `functional::` (reference include/functional.h:174,180;
 * src/device.cu and include/device.cuh are empty).  Every entry point below names the reference code whose
 * results it reproduces (file:line relative to the reference root).  INTEGRATION.md shows the binding a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; gnn_last_error() gives the message
 *     (thread-local).  The C++ host shim rethrows std::runtime_error like the reference's CHECK_* helpers
 *     (reference include/utils.h:19-30).
 *   - pointers are DEVICE pointers unless the parameter name ends in `_h` (host).
 *   - all kernels are enqueued on the context's stream; nothing synchronises unless stated.
 *   - matrices are dense row-major fp32 with an explicit leading dimension (elements).
 *   - there is no CPU fallback: creating a context without a CUDA device fails.
 */
#ifndef GNN_C_H
#define GNN_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define GNN_API __attribute__((visibility("default")))
#else
#define GNN_API
#endif

typedef struct gnn_ctx gnn_ctx_t;
typedef struct gnn_graph gnn_graph_t;
typedef struct gnn_gcn gnn_gcn_t;
typedef struct gnn_peer_arena gnn_peer_arena_t;

/* ---------------------------------------------------------------- context / errors ---------------- */
GNN_API int gnn_version(void);
GNN_API const char *gnn_last_error(void);
/* stream: an existing cudaStream_t to enqueue on (e.g. the caller's), or NULL to create one. */
GNN_API int gnn_ctx_create(int device, void *stream, gnn_ctx_t **out);
GNN_API int gnn_ctx_destroy(gnn_ctx_t *ctx);
GNN_API int gnn_ctx_sync(gnn_ctx_t *ctx);
GNN_API void *gnn_ctx_stream(gnn_ctx_t *ctx);
GNN_API int gnn_ctx_sm_count(gnn_ctx_t *ctx);
/* number of kernels this library has launched on ctx since creation (bench.py's gpu_launches). */
GNN_API int64_t gnn_ctx_launch_count(gnn_ctx_t *ctx);

/* ---------------------------------------------------------------- storage -------------------------
 * Replaces the heap std::valarray storage of cyg::tensor (reference include/tensor.h:825-828).  gnn_free waits for the
 * context's stream before releasing the block. */
GNN_API int gnn_malloc(gnn_ctx_t *ctx, void **ptr, size_t bytes);
GNN_API int gnn_free(gnn_ctx_t *ctx, void *ptr);
GNN_API int gnn_memset(gnn_ctx_t *ctx, void *ptr, int value, size_t bytes);
GNN_API int gnn_memcpy_h2d(gnn_ctx_t *ctx, void *dst, const void *src_h, size_t bytes); /* async on stream */
GNN_API int gnn_memcpy_d2h(gnn_ctx_t *ctx, void *dst_h, const void *src, size_t bytes); /* synchronises */
GNN_API int gnn_memcpy_d2d(gnn_ctx_t *ctx, void *dst, const void *src, size_t bytes);
GNN_API int gnn_fill_f32(gnn_ctx_t *ctx, float *ptr, float value, int64_t n);

/* ---------------------------------------------------------------- graph structure (K1-K3) ----------
 * gnn_graph_build: COO edge list -> CSR of the 0/1 adjacency, rows ascending, columns ascending and
 * de-duplicated inside a row.  Reproduces, without the dense N x N round trip,
 *     graph::edge_to_adj_mat      reference src/graph.cpp:21-44   (A[src][dst] = 1, duplicates collapse)
 *     tensor::fill_diagonal_      reference include/tensor.h:806-817
 *     graph::adj_to_edge_list     reference src/graph.cpp:46-67   (row-major scan -> sorted COO)
 *     graph::add_self_loops       reference src/graph.cpp:68-75
 * fill_mode: 0 = diagonal removed (add_self_loops fillValue 0, as GCNConv calls it, src/graph.cpp:172),
 *            1 = diagonal forced on every row (fillValue 1; the A+I of the north-star layer),
 *            2 = diagonal left as given.
 * src/dst are device int32 arrays of length E (row = src = edge_index[0], col = dst = edge_index[1]). */
GNN_API int gnn_graph_build(gnn_ctx_t *ctx, const int32_t *src, const int32_t *dst, int64_t E, int32_t N,
                            int fill_mode, gnn_graph_t **out);
GNN_API int gnn_graph_build_h(gnn_ctx_t *ctx, const int32_t *src_h, const int32_t *dst_h, int64_t E, int32_t N,
                              int fill_mode, gnn_graph_t **out);
/* Weighted adjacency — graph::edge_to_adj_mat with edge_attr (reference src/graph.cpp:21-44): A[src][dst] = w by
 * assignment in edge order, so the LAST weight of a duplicated (src, dst) wins; fill_mode 1 then forces the diagonal to 1
 * (tensor::fill_diagonal_(1): the A + I of the north-star layer), 2 leaves it as given.  w: device float[E].
 * The structure equals the unweighted build; gnn_graph_normalize then uses the weighted degree
 * deg = rowsum(A), dinv = deg^-1/2, val = (A * dinv) * dinv^T. */
GNN_API int gnn_graph_build_weighted(gnn_ctx_t *ctx, const int32_t *src, const int32_t *dst, const float *w, int64_t E,
                                     int32_t N, int fill_mode, gnn_graph_t **out);
/* raw weights of the stored entries (host float[nnz], CSR order); synchronises */
GNN_API int gnn_graph_export_weights_h(gnn_ctx_t *ctx, const gnn_graph_t *g, float *val0_h);
/* Adopt an existing device CSR (used by the row-partitioned multi-GPU path: local rows, global columns).
 * n_rows x n_cols, rowptr[n_rows+1] int32, colidx[nnz] int32, val[nnz] fp32 or NULL; arrays are copied. */
GNN_API int gnn_graph_from_csr(gnn_ctx_t *ctx, int32_t n_rows, int32_t n_cols, const int32_t *rowptr,
                               const int32_t *colidx, const float *val, gnn_graph_t **out);
/* K2: CSC (= CSR of the transpose) + permutation, rows ascending inside a column.  The reference instead
 * clones and transposes the dense matrix on every backward (include/operation.h:526-528). */
GNN_API int gnn_graph_build_csc(gnn_ctx_t *ctx, gnn_graph_t *g);
/* K3: deg = rowsum(A0+I) (src/graph.cpp:178), dinv = deg^-1/2 (src/graph.cpp:183),
 * val[r,c] = dinv[r]*dinv[c] (mode-B composition of functional::mul, include/functional.h:189-213);
 * also valT when the CSC exists. */
GNN_API int gnn_graph_normalize(gnn_ctx_t *ctx, gnn_graph_t *g);
/* Normalisation of graph::GCNConv::forward AS WRITTEN (reference src/graph.cpp:176-185) for a graph built with
 * fill_mode 0 (its add_self_loops(.., 0) removes the loops): deg = rowsum(A0)+1, dinv = deg^-1/2,
 * norm = (A0 dinv) * dinv, and the stored values become val[r,c] = norm[r], so that gnn_spmm_fwd(use_values=1) computes
 * aggregate_and_update's (A0 h) * norm (src/graph.cpp:204-212) and gnn_spmm_bwd its transpose.  norm_out: optional
 * device float[N]. */
GNN_API int gnn_graph_normalize_as_written(gnn_ctx_t *ctx, gnn_graph_t *g, float *norm_out);
GNN_API int gnn_graph_destroy(gnn_ctx_t *ctx, gnn_graph_t *g);
GNN_API int64_t gnn_graph_nnz(const gnn_graph_t *g);
GNN_API int32_t gnn_graph_rows(const gnn_graph_t *g);
GNN_API int32_t gnn_graph_cols(const gnn_graph_t *g);
/* 1 when the CSC arrays alias the CSR arrays (structurally symmetric graph, detected on device). */
GNN_API int gnn_graph_is_symmetric(const gnn_graph_t *g);
/* Copy structure arrays to host for bit-exact checks (any pointer may be NULL).  Sizes: rowptr/colptr
 * rows+1 / cols+1 int32, colidx/rowidx/perm/val/valT nnz, deg/dinv rows. Synchronises. */
GNN_API int gnn_graph_export_h(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t *rowptr_h, int32_t *colidx_h,
                               float *val_h, int32_t *colptr_h, int32_t *rowidx_h, int32_t *perm_h, float *valT_h,
                               int32_t *deg_h, float *dinv_h);
/* Dense adjacency (graph::Data::to_adj / edge_to_adj_mat, src/graph.cpp:118-129) for API fidelity at small N:
 * out[n_rows, n_cols] = 1 (weighted 0), the normalised value (1) or the raw edge weight (2: Data::to_adj with edge_attr)
 * at the stored positions, 0 elsewhere. */
GNN_API int gnn_graph_to_dense(gnn_ctx_t *ctx, const gnn_graph_t *g, int weighted, float *out, int64_t ld);
/* Dense matrix -> row-major sorted COO of the entries with int(a) != 0 (graph::adj_to_edge_list,
 * src/graph.cpp:46-67).  Two passes (flag + scan + compact); *count_h receives the number of entries; when
 * out_rows/out_cols are NULL only the count is produced. Synchronises. */
GNN_API int gnn_dense_to_coo(gnn_ctx_t *ctx, const float *A, int64_t rows, int64_t cols, int64_t ld,
                             int32_t *out_rows, int32_t *out_cols, float *out_vals, int64_t capacity,
                             int64_t *count_h);

/* ---------------------------------------------------------------- aggregation (K4/K5/K7) -----------
 * Forward  Y[n_rows,F] = A_hat * P           — adj_mat->mm(x), reference src/graph.cpp:208 ->
 *                                              include/functional.h:398-441
 * Backward dP[n_cols,F] = A_hat^T * dZ       — MatMul::_backward rhs branch, include/operation.h:524-531,
 *                                              computed over the CSC: no atomics, fixed order.
 * Fused epilogue (all optional):
 *     bias[F]        Z = Y + b               — Add, include/operation.h:102-129
 *     relu           H = Z > 0 ? Z : 0       — nn::ReLU / Mask, src/nn.cpp:229-237, operation.h:537-573
 *     mask[.,F]      out = mask > 0 ? out : 0 (ReLU backward, operation.h:557-562)
 * use_values = 0 treats every stored entry as 1 (plain sum aggregation, graph.cpp:204-212).
 * Padding contract: columns >= F of a row are never written.  The 128-bit kernels are used when P and Y are 16-byte
 * aligned with leading dimensions that are multiples of 4 and either F % 4 == 0 or both leading dimensions equal
 * round_up(F, 4) (then the padding columns F..ld-1 of P are read and those of Y are overwritten with the aggregate of
 * P's padding); every other view — e.g. a 47-column slice of a 256-wide matrix — takes the scalar kernels. */
/* which aggregation kernel the automatic choice takes for this structure (1 = one row per lane group, 2 = nonzero-balanced
 * merge kernel): by degree skew, longest row >= 16 x the mean row length -> 2 (see csrc/spmm.cu use_merge). */
GNN_API int gnn_graph_spmm_variant(gnn_ctx_t *ctx, const gnn_graph_t *g, int transpose);
GNN_API int gnn_spmm_fwd(gnn_ctx_t *ctx, const gnn_graph_t *g, const float *P, int64_t ldp, int32_t F, float *Y,
                         int64_t ldy, const float *bias, int relu, const float *mask, int64_t ldm, int use_values);
GNN_API int gnn_spmm_bwd(gnn_ctx_t *ctx, const gnn_graph_t *g, const float *dZ, int64_t ldz, int32_t F, float *dP,
                         int64_t ldp, const float *mask, int64_t ldm, int use_values);
/* SpMM variant selection: 0 = auto (the nonzero-balanced kernel unless the matrix has an empty row), 1 = one output
 * row per lane group, 2 = nonzero-balanced ("merge"): equal chunks of nonzeros per warp, rows cut by a chunk boundary
 * are completed by a fixed-order fix-up pass (no atomics).  Higher decimal digits are tuning knobs for experiments
 * (csrc/spmm.cu). */
GNN_API int gnn_set_spmm_variant(gnn_ctx_t *ctx, int variant);

/* ---------------------------------------------------------------- dense transforms (K6) ------------
 *   gnn_gemm_nt: C[M,N] = A[M,K] * B[N,K]^T (+bias[N]) (relu)   — nn::Linear::forward, src/nn.cpp:205-211
 *   gnn_gemm_nn: C[M,N] = A[M,K] * B[K,N] (mask)                — MatMul::_backward lhs branch, operation.h:516-523
 *   gnn_gemm_tn: C[K1,K2] = A[M,K1]^T * B[M,K2]                 — rhs branch + Transpose::_backward,
 *                                                                  operation.h:524-531,416-433 (fixed-order split over M)
 * precision: 0 = FP32 FMA (CUDA cores), 1 = 3xTF32 on tcgen05 tensor cores (error ~2^-21, inside 1e-5); wide products
 * (N >= 128 with K > 64; K1 > 128 for gnn_gemm_tn) run on CTA pairs (tcgen05.mma.cta_group::2, clusters of 2), with
 * results bit-identical to the single-CTA kernels.  Environment switches for A/B runs: GNN_GEMM_PAIR (bit 0 NT/NN,
 * bit 1 TN; default 3), GNN_FUSED_BIAS_GRAD=0 (fused trainer: separate bias-gradient kernel), GNN_SPMM_ASYNC=8
 * (read at gnn_ctx_create: aggregation gathers staged through shared memory, an ablation). */
GNN_API int gnn_gemm_nt(gnn_ctx_t *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B,
                        int64_t ldb, float *C, int64_t ldc, const float *bias, int relu, int precision);
GNN_API int gnn_gemm_nn(gnn_ctx_t *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B,
                        int64_t ldb, float *C, int64_t ldc, const float *mask, int64_t ldm, int precision);
GNN_API int gnn_gemm_tn(gnn_ctx_t *ctx, int64_t M, int32_t K1, int32_t K2, const float *A, int64_t lda,
                        const float *B, int64_t ldb, float *C, int64_t ldc, int precision);

/* ---------------------------------------------------------------- epilogues, loss, optimiser -------
 * bias+ReLU forward / ReLU mask backward / bias gradient (ascending-row column sum; Add::_backward ->
 * sum_to_size, include/tensor.h:618-638). */
GNN_API int gnn_bias_relu_fwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *Y, int64_t ldy, const float *bias,
                              int relu, float *out, int64_t ldo);
GNN_API int gnn_relu_bwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *dH, int64_t ldd, const float *act,
                         int64_t lda, float *dZ, int64_t ldo);
GNN_API int gnn_bias_grad(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *dZ, int64_t ldd, float *db);
/* Fused softmax-cross-entropy: loss (device scalar) = mean_i -log(exp(z_iy)/(sum_c exp(z_ic)+1e-20))
 * (nn::cross_entropy_loss forward, src/nn.cpp:442-453); dZ = (softmax(Z)-onehot(y))/N_total (the analytic
 * gradient — the reference backward throws, SURVEY.md bug B3).  n_total is the divisor (global node count
 * under row partitioning); dZ may be NULL. */
GNN_API int gnn_softmax_xent(gnn_ctx_t *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                             int64_t n_total, float *loss, float *dZ, int64_t ldd);
/* torch.optim.SGD semantics, the documented intent of nn::SGD (include/nn.h:165-178; body broken, bug B4).
 * vel may be NULL when momentum == 0; first = 1 on the first step (velocity initialised to the gradient). */
GNN_API int gnn_sgd_step(gnn_ctx_t *ctx, int64_t n, float *p, const float *g, float *vel, float lr, float momentum,
                         float dampening, float weight_decay, int nesterov, int first);

/* nn::BatchNorm over the node dimension with training statistics (reference src/nn.cpp:301-330): per feature
 * mean = sum x / N, var = sum (x-mean)^2 / N (two passes, functional::var with correction 0),
 * Y = (X - mean) / sqrt(var + eps) * gamma (+ beta) (ReLU optional: GCNConv applies nn::ReLU right after).
 * mean/var (device float[F]) are outputs, kept by the caller for the backward and the running statistics.
 * Backward: the standard batch-norm gradient with the ReLU mask taken from relu_out (the forward output; NULL = no
 * ReLU); dgamma/dbeta are device float[F].  (The reference's own autograd loses fan-out gradients here, bug B2.) */
GNN_API int gnn_batchnorm_fwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *gamma,
                              const float *beta, float eps, int relu, float *Y, int64_t ldy, float *mean, float *var);
GNN_API int gnn_batchnorm_bwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *mean,
                              const float *var, const float *gamma, float eps, const float *relu_out, int64_t ldy,
                              const float *dY, int64_t ldd, float *dX, int64_t ldo, float *dgamma, float *dbeta);
/* nn::LayerNorm over the feature dimension (reference src/nn.cpp:332-353): per row mean / biased variance (two passes),
 * Y = (X - mean) / sqrt(var + eps) * gamma + beta (gamma/beta may be NULL), optional fused ReLU (nn::MLP applies nn::ReLU
 * next, include/nn.h:193-214); mean / rstd are device float[N] outputs for the backward.  Backward: the standard
 * layer-norm gradient, ReLU mask from relu_out (the forward output, NULL = none); dgamma/dbeta device float[F] or NULL. */
GNN_API int gnn_layernorm_fwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *gamma,
                              const float *beta, float eps, int relu, float *Y, int64_t ldy, float *mean, float *rstd);
GNN_API int gnn_layernorm_bwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *mean,
                              const float *rstd, const float *gamma, const float *relu_out, int64_t ldy, const float *dY,
                              int64_t ldd, float *dX, int64_t ldo, float *dgamma, float *dbeta);
/* nn::tanh (src/nn.cpp:355-364; evaluated with tanhf, i.e. without the overflow of the as-written exp quotient) and its
 * gradient dx = dy * (1 - y^2) */
GNN_API int gnn_tanh_fwd(gnn_ctx_t *ctx, int64_t n, const float *x, float *y);
GNN_API int gnn_tanh_bwd(gnn_ctx_t *ctx, int64_t n, const float *y, const float *dy, float *dx);
/* nn::Dropout (src/nn.cpp:246-266): y = keep ? x / (1 - p) : 0 with a counter-based keep mask that depends only on
 * (seed, element index) — the reference seeds a fresh engine from time() on every call (bug B6).  The backward is the
 * same call on the incoming gradient with the same seed. */
GNN_API int gnn_dropout(gnn_ctx_t *ctx, int64_t n, const float *x, float p, uint64_t seed, float *y);
/* torch.optim.Adam semantics — the intent of nn::Adam (include/nn.h:180-188); the reference body (src/nn.cpp:419-441)
 * divides by sqrt(v)*eps and uses the parameter index as step count.  m, v: first/second moment buffers (zeroed by
 * the caller before step 1); step counts from 1. */
GNN_API int gnn_adam_step(gnn_ctx_t *ctx, int64_t n, float *p, const float *g, float *m, float *v, float lr, float beta1,
                          float beta2, float eps, float weight_decay, int64_t step);
/* Loss over the rows selected by a node mask (graph::Data::set_mask train/val/test masks, src/graph.cpp:130-151):
 * mean over n_selected of the same per-row term as gnn_softmax_xent; dZ rows of unselected nodes are zero.
 * mask: device uint8[N] (the reference's tensor<bool>). */
GNN_API int gnn_softmax_xent_masked(gnn_ctx_t *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                                    const uint8_t *mask, int64_t n_selected, float *loss, float *dZ, int64_t ldd);
/* Number of rows (selected by mask, or all when mask is NULL) whose arg-max logit (first maximum, tensor::argmax,
 * include/tensor.h:645-648) equals the label; *count is a device int64. */
GNN_API int gnn_argmax_correct(gnn_ctx_t *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                               const uint8_t *mask, int64_t *count);

/* ---------------------------------------------------------------- elementwise / reductions for the
 * cyg::tensor surface (functional::add/mul/div/exp/log/sum/transpose/mask, include/functional.h:162-471) */
enum { GNN_OP_ADD = 0, GNN_OP_MUL = 1, GNN_OP_DIV = 2, GNN_OP_POW = 3, GNN_OP_GT = 4 };
enum { GNN_UOP_EXP = 0, GNN_UOP_LOG = 1, GNN_UOP_NEG = 2 };
/* out[rows,cols] = a (op) b with numpy broadcasting expressed as strides (0 = broadcast) */
GNN_API int gnn_binary_f32(gnn_ctx_t *ctx, int op, int64_t rows, int64_t cols, const float *a, int64_t a_rs,
                           int64_t a_cs, const float *b, int64_t b_rs, int64_t b_cs, float *out);
GNN_API int gnn_unary_f32(gnn_ctx_t *ctx, int op, int64_t n, const float *a, float *out);
/* out = cond > 0 ? t : f  (functional::mask, include/functional.h:443-471) */
GNN_API int gnn_where_f32(gnn_ctx_t *ctx, int64_t n, const float *cond, const float *t, const float *f, float *out);
/* sum over dim of a [rows, cols] view: dim 0 -> out[cols], dim 1 -> out[rows], dim -1 (all) -> out[1] */
GNN_API int gnn_sum_f32(gnn_ctx_t *ctx, int64_t rows, int64_t cols, const float *a, int dim, float *out);
GNN_API int gnn_transpose_f32(gnn_ctx_t *ctx, int64_t rows, int64_t cols, const float *a, float *out);
/* out[i] = a[i, idx[i]]  (tensor::at / functional::slice, include/functional.h:482-494) */
GNN_API int gnn_gather_cols_f32(gnn_ctx_t *ctx, int64_t rows, int64_t cols, const float *a, const int32_t *idx,
                                float *out);

/* ---------------------------------------------------------------- fused trainer --------------------
 * One full-batch GCN train step (forward + loss + backward + SGD) for the layer stack
 *     Z_l = A_hat (H_{l-1} W_l^T) + b_l,  H_l = ReLU(Z_l) (l < L),  logits = Z_L
 * i.e. what a main.cpp loop over graph::GCNConv / nn::cross_entropy_loss / nn::SGD executes,
 * with every buffer preallocated so the step is a fixed launch sequence (CUDA-graph capturable).
 * Per layer the aggregation runs at the narrower of (F_{l-1}, F_l): A_hat(HW^T) == (A_hat H)W^T.
 * The graph must have CSC + normalisation built.  dims has L+1 entries. */
GNN_API int gnn_gcn_create(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t L, const int32_t *dims, gnn_gcn_t **out);
GNN_API int gnn_gcn_destroy(gnn_ctx_t *ctx, gnn_gcn_t *m);
/* layer = 1..L.  W[F_l, F_{l-1}] row-major (nn::Linear layout, src/nn.cpp:187-194), b[F_l]. */
GNN_API int gnn_gcn_set_params_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, const float *W_h, const float *b_h);
GNN_API int gnn_gcn_get_params_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, float *W_h, float *b_h);
GNN_API int gnn_gcn_get_grads_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, float *dW_h, float *db_h);
/* activations of layer l (pre-ReLU Z_l is not kept for l < L; this returns H_l = ReLU(Z_l), logits for l = L) */
GNN_API int gnn_gcn_get_activation_h(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, float *out_h);
GNN_API int gnn_gcn_get_dlogits_h(gnn_ctx_t *ctx, gnn_gcn_t *m, float *out_h);
/* options: precision (0 fp32 / 1 3xTF32), agg_first_mask, profile, momentum, dampening, weight_decay, nesterov,
 * optimizer (0 SGD / 1 Adam), beta1, beta2, eps */
GNN_API int gnn_gcn_set_option(gnn_gcn_t *m, const char *key, double value);
/* X[N,F0] (ld = ldx) and y[N] on the device.  loss_d: device float written by the step. */
GNN_API int gnn_gcn_train_step(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X, int64_t ldx, const int32_t *y,
                               float lr, float *loss_d);
/* forward only (inference); logits stay in the model's activation buffer */
GNN_API int gnn_gcn_forward(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X, int64_t ldx);
/* End-to-end step from HOST buffers: copies X_h and y_h to the device (pinned or pageable), runs the step,
 * copies the loss back and synchronises.  This is the call bench.py's `e2e` times. */
GNN_API int gnn_gcn_train_step_h(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X_h, const int32_t *y_h, float lr,
                                 float *loss_h);
/* Enqueue the upload of the NEXT step's host inputs on a copy stream (double-buffered device slots), so it
 * overlaps the current step's kernels.  The following gnn_gcn_train_step_h runs on the prefetched batch; the
 * X_h / y_h it is given are then taken as the batch AFTER that one and prefetched in turn (pass NULL to stop
 * pipelining).  So a loop `prefetch(b0); for k: train_step_h(b[k+1])` uploads every step's inputs exactly once,
 * hidden behind the previous step.  Host buffers must be pinned and stay valid until consumed. */
GNN_API int gnn_gcn_prefetch_h(gnn_ctx_t *ctx, gnn_gcn_t *m, const float *X_h, const int32_t *y_h);
/* Training-node mask (graph::Data::set_mask(mask, TRAIN), src/graph.cpp:130-151): device uint8[local rows]; the
 * loss becomes the mean over the n_selected_total selected nodes of the whole graph (all ranks).  NULL = all nodes. */
/* ReLU tie-break overrides for cross-implementation parity checks: entries (rows[i], cols[i]) — LOCAL row ids — of
 * hidden layer `layer`'s output H_l are forced, right after the forward produces them, to max(H, 1e-30) (positive[i] != 0)
 * or to 0.  `Z > 0` (nn::ReLU, src/nn.cpp:229-237; Mask::_backward, operation.h:557-562) is discontinuous, so two correct
 * fp32 implementations legitimately disagree on the mask of pre-activations within rounding distance of zero; listing
 * exactly those entries (|Z| <= 1e-5 max|Z| in the other implementation) makes the backward masks identical while
 * changing no forward value by more than the 1e-5 tolerance.  Device arrays, caller-owned; n = 0 clears the layer. */
GNN_API int gnn_gcn_set_relu_overrides(gnn_ctx_t *ctx, gnn_gcn_t *m, int32_t layer, const int32_t *rows, const int32_t *cols,
                                       const uint8_t *positive, int64_t n);
GNN_API int gnn_gcn_set_train_mask(gnn_ctx_t *ctx, gnn_gcn_t *m, const uint8_t *mask, int64_t n_selected_total);
/* correct predictions (arg-max of the logits of the last forward) among the rows selected by mask (val/test mask;
 * NULL = all local rows); *count is a device int64 (per rank: sum over ranks on the host side). */
GNN_API int gnn_gcn_accuracy(gnn_ctx_t *ctx, gnn_gcn_t *m, const int32_t *y, const uint8_t *mask, int64_t *count);
/* per-kernel-class device time of the last profiled step, ms: fills up to n entries of
 * {spmm, gemm, loss, bias_grad, sgd, other}.  Enabled by option "profile" = 1 (adds event records). */
GNN_API int gnn_gcn_last_breakdown(gnn_gcn_t *m, double *ms, int n);
/* every aggregation launch of the last profiled step, in execution order: device time (ms), algorithmic bytes
 * (4(rows+1) + nnz(8+4F) + 4 rows F) and width F; *n = number of launches (may exceed cap) */
GNN_API int gnn_gcn_last_spmm_spans(gnn_gcn_t *m, double *ms, double *alg_bytes, int32_t *F, int cap, int *n);
/* algorithmic SpMM bytes / number of SpMM launches in one train step (for the roofline line) */
GNN_API int gnn_gcn_spmm_stats(gnn_gcn_t *m, double *alg_bytes, int32_t *n_spmm, double *gemm_flops);
/* how the aggregation inputs of this model travel between ranks (what actually runs, for the bench line):
 * 0 single GPU, 1 ncclAllGather per aggregation, 2 peer-arena pushes by the SM store kernel (default),
 * 3 peer-arena pushes by copy engines, 4 pipelined panels with in-place ncclAllGather as transport,
 * 6 2-D partition (gnn_gcn_create_grid: column-slice scatter before, fused row exchange inside every aggregation) */
GNN_API int gnn_gcn_exchange_mode(const gnn_gcn_t *m);
/* 2-D partition only.  halo_fraction: rows of this rank its peers' structure blocks read / (peers x local rows) — 1 means
 * every peer needs every row (the synthetic power-law graphs), small values mean the graph has locality; halo_lists: 1 when
 * the exchange sends only those rows (chosen when halo_fraction < 0.9); split: 1 when the aggregation runs as interior rows
 * (only own rows needed, overlapping the exchange) + boundary rows; interior_fraction: interior rows / structure rows. */
GNN_API int gnn_gcn_exchange_stats(const gnn_gcn_t *m, double *halo_fraction, int *halo_lists, int *split, double *interior_fraction);
/* Fused trainer under a 2-D partition of the aggregation: world = Pr row groups x Pc feature-column groups, this rank =
 * gi * Pc + gj.  `g` holds the structure rows of row group gi — gnn_graph_slice_rows(global, lo, hi) with
 * lo = min(N, gi Pc c), hi = min(N, (gi+1) Pc c), c = ceil(N / world) — all columns.  Activations, X and y stay 1-D
 * row-partitioned (rank r owns rows [r c, (r+1) c)), so every other gnn_gcn_* entry point is used unchanged.
 * Per aggregation each rank receives N F / Pc + c F (Pc-1)/Pc floats instead of the N F (world-1)/world of the 1-D row
 * partition's all-gather (src/graph.cpp:204-212 is the aggregation being sharded).  Needs CUDA IPC peer mapping
 * (returns 5 otherwise; there is no NCCL variant of this mode). */
/* Host-only rules of that partition (what the trainer itself uses; exported for plans and CPU tests):
 * column slice [c0, c0 + w) that column group j of Pc receives of a matrix of padded width ldw (16-byte units dealt out
 * in order, the first ldw/4 % Pc groups get one more); and for `rank` of `world` = Pr x Pc the activation rows
 * [rows_lo, rows_hi) it owns and the structure rows [group_lo, group_hi) of its row group (rank / Pc). */
GNN_API int gnn_partition_col_slice_h(int32_t ldw, int32_t Pc, int32_t j, int32_t *c0_h, int32_t *w_h);
GNN_API int gnn_partition_grid_h(int64_t N, int32_t world, int32_t Pc, int32_t rank, int64_t *rows_lo_h, int64_t *rows_hi_h,
                                 int64_t *group_lo_h, int64_t *group_hi_h);
GNN_API int gnn_gcn_create_grid(gnn_ctx_t *ctx, const gnn_graph_t *g, int32_t L, const int32_t *dims, int32_t Pr, int32_t Pc,
                                gnn_gcn_t **out);

/* Measured dense TF32 tensor-core peak of the device (TFLOP/s, 2MNK per 128 x 256 x 8 tcgen05.mma kind::tf32 streamed
 * from shared-memory operands on every SM): the compute-side roofline denominator of the dense transforms, which issue
 * three such MMAs per product (3xTF32). */
GNN_API int gnn_tf32_peak_probe(gnn_ctx_t *ctx, double *tflops_h);

/* ---------------------------------------------------------------- multi-GPU (K10) ------------------
 * 1-D contiguous row partition: part_ptr[p] = min(N, p*ceil(N/P)).  Each rank owns the CSR rows (and CSC
 * columns) of its nodes with GLOBAL column ids; per aggregation the ranks all-gather their feature row
 * blocks (NCCL over NVLink) into a global-order buffer and aggregate locally. */
GNN_API int gnn_partition_ptr_h(int64_t N, int32_t P, int64_t *part_ptr_h);
/* Column panels of a gathered matrix of padded width ldw (multiple of 4): at most 4 panels of panel_cols columns
 * (wider when ldw > 4 * panel_cols), starts c0_h[p] and widths w_h[p] (multiples of 4), *n_h panels.  Panel p of a
 * gather region is stored panel-major, [world][rows_per_rank, w_h[p]], so a rank's panel is one contiguous range.
 * Host-only (no device needed): the rule the fused trainer uses, exported for plans and tests. */
GNN_API int gnn_partition_panels_h(int32_t ldw, int32_t panel_cols, int32_t *c0_h, int32_t *w_h, int32_t *n_h);
/* Extract rank-local rows [lo,hi) of g as a new graph (n_rows = hi-lo, n_cols = N, global column ids),
 * including values and, when g has them, the matching CSC slice for the backward. */
GNN_API int gnn_graph_slice_rows(gnn_ctx_t *ctx, const gnn_graph_t *g, int64_t lo, int64_t hi, gnn_graph_t **out);
/* Per-rank partition arrays of a row block, built on the device, BIT-EXACT against the CPU
 * restatement (oracle: orc_partition_halo / orc_partition_interior).  `g` is a gnn_graph_slice_rows block holding rows
 * [lo, hi) with global column ids; transpose = 1 uses its backward (CSC) block instead.
 *   halo_ids[n_halo]      sorted unique columns outside [lo, hi): the remote feature rows the block needs
 *   local_colidx[nnz]     columns renumbered for a [own rows ; halo rows] input matrix (c - lo, or n_loc + halo position)
 *   interior[n_rows]      1 = every column is owned: the row can be aggregated before any halo row has arrived;
 *   interior_rows / boundary_rows: the two ascending row lists (overlap of the exchange with interior-row aggregation).
 * Export pointers may be NULL. */
typedef struct gnn_partition gnn_partition_t;
GNN_API int gnn_partition_build(gnn_ctx_t *ctx, const gnn_graph_t *g, int64_t lo, int64_t hi, int transpose,
                                gnn_partition_t **out);
GNN_API int64_t gnn_partition_halo_count(const gnn_partition_t *p);
GNN_API int64_t gnn_partition_interior_count(const gnn_partition_t *p);
GNN_API int64_t gnn_partition_nnz(const gnn_partition_t *p);
GNN_API int gnn_partition_export_h(gnn_ctx_t *ctx, const gnn_partition_t *p, int32_t *halo_ids_h, int32_t *local_colidx_h,
                                   uint8_t *interior_h, int32_t *interior_rows_h, int32_t *boundary_rows_h);
GNN_API int gnn_partition_destroy(gnn_ctx_t *ctx, gnn_partition_t *p);
/* NCCL communicator owned by the context.  id_h: 128-byte ncclUniqueId produced by gnn_comm_unique_id_h on
 * rank 0 and distributed by the caller (torch.distributed / MPI / files). */
GNN_API int gnn_comm_unique_id_h(void *id_h /* 128 bytes */);
GNN_API int gnn_comm_init(gnn_ctx_t *ctx, const void *id_h, int rank, int world);
GNN_API int gnn_comm_destroy(gnn_ctx_t *ctx);
/* all-gather equal row blocks: recv[world*rows_per_rank, F] <- send[rows_per_rank, F] (dense, ld = F) */
GNN_API int gnn_allgather_rows(gnn_ctx_t *ctx, const float *send, float *recv, int64_t rows_per_rank, int32_t F);
GNN_API int gnn_allreduce_sum(gnn_ctx_t *ctx, float *buf, int64_t n);
/* Peer arena: the exchange the fused trainer uses instead of ncclAllGather.  Every rank allocates `bytes` of
 * device memory, the ranks swap CUDA IPC handles through the communicator and map each other's arenas (collective
 * call; returns 5 when IPC peer mapping is unavailable and the caller stays on gnn_allgather_rows).
 * Ranks exchange byte ranges of their arenas: a rank produces data in place in its own arena, then
 *   gnn_peer_gather_begin: after the work already enqueued on the context's stream, push [offset, offset+bytes) of
 *       the own arena to the same offset of every peer's arena over NVLink (SM store kernel on a high-priority side
 *       stream; GNN_PEER_COPY=ce selects copy engines) and publish the slot's next sequence number in every peer's
 *       flag word for (slot, this rank);
 *   gnn_peer_gather_wait:  make the context's stream wait until, for each of the `count` slots starting at `slot`,
 *       every peer's latest push on that slot has landed.  Work enqueued between the two calls overlaps the
 *       transfer.  Every rank must issue the same sequence of begin calls per slot.
 * A range may be rewritten only after a collective that orders all ranks (the trainer's gradient all-reduce) —
 * the trainer gives every aggregation of a step its own slots and region. */
GNN_API int gnn_peer_arena_create(gnn_ctx_t *ctx, size_t bytes, gnn_peer_arena_t **out);
GNN_API int gnn_peer_arena_destroy(gnn_ctx_t *ctx, gnn_peer_arena_t *a);
GNN_API void *gnn_peer_arena_local(gnn_peer_arena_t *a);
GNN_API int gnn_peer_gather_begin(gnn_ctx_t *ctx, gnn_peer_arena_t *a, int slot, size_t offset, size_t bytes);
GNN_API int gnn_peer_gather_wait(gnn_ctx_t *ctx, gnn_peer_arena_t *a, int slot, int count);

#ifdef __cplusplus
}
#endif
#endif /* GNN_C_H */
