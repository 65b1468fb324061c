// spmm.cu — aggregation kernels K4/K5:  Y = A_hat * P  over a CSR (forward) or over the CSC of A_hat
// (= CSR of A_hat^T, backward), with the bias / ReLU / ReLU-mask epilogue of K7 fused in.
//
// HBM/L2-gather bound.  Design (B200):
//   - a group of LPR lanes owns one output row; lanes hold VEC vectors (128-bit when the leading
//     dimensions allow) of the F-wide accumulator in registers, so each output row is written once and
//     there are NO atomics: the per-row order is the stored (ascending-column) order, i.e. deterministic;
//   - the group's lanes load LPR (column, value) pairs at a time, coalesced, and broadcast them with
//     shuffles; feature rows are gathered with U independent 128-bit loads in flight per lane;
//   - index/value streams use the no-allocate path (read once), feature rows the default path (L2 reuse);
//   - rows are distributed over a 1-D grid sized to whole waves of 148 SMs x resident CTAs.
#include "common.cuh"

namespace gnn {

template <typename V> struct VecTraits;
template <> struct VecTraits<float4> {
    static constexpr int W = 4;
    static __device__ __forceinline__ float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ void fma(float4 &acc, float a, const float4 &p) {
        acc.x = fmaf(a, p.x, acc.x); acc.y = fmaf(a, p.y, acc.y);
        acc.z = fmaf(a, p.z, acc.z); acc.w = fmaf(a, p.w, acc.w);
    }
    static __device__ __forceinline__ void add(float4 &acc, const float4 &p) {
        acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
};
template <> struct VecTraits<float> {
    static constexpr int W = 1;
    static __device__ __forceinline__ float zero() { return 0.f; }
    static __device__ __forceinline__ void fma(float &acc, float a, const float &p) { acc = fmaf(a, p, acc); }
    static __device__ __forceinline__ void add(float &acc, const float &p) { acc += p; }
};

__device__ __forceinline__ int32_t ld_stream_i32(const int32_t *p) {
    int32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ float epilogue1(float v, int col, int F, const float *__restrict__ bias, int relu,
                                           const float *__restrict__ mrow) {
    if (col < F) {
        if (bias) v += bias[col];
        if (relu) v = v > 0.f ? v : 0.f;            // NaN -> 0 like functional::mask (functional.h:460-461)
        if (mrow) v = mrow[col] > 0.f ? v : 0.f;    // Mask::_backward (operation.h:557-562)
    }
    return v;
}

constexpr int SPMM_THREADS = 256;
constexpr int SPMM_U = 4; // independent feature-row loads in flight per lane and vector slot

// V = float4: requires P/Y 16-byte aligned and ldp/ldy multiples of 4 (padding columns may be touched).
// V = float : no alignment requirement.
template <typename V, int LPR, int VEC, bool USE_VAL>
__global__ void __launch_bounds__(SPMM_THREADS)
    spmm_rows_kernel(int32_t n_out, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                     const float *__restrict__ val, const float *__restrict__ P, int64_t ldp, int32_t F,
                     float *__restrict__ Y, int64_t ldy, const float *__restrict__ bias, int relu,
                     const float *__restrict__ mask, int64_t ldm) {
    using T = VecTraits<V>;
    constexpr int W = T::W;
    constexpr int GROUPS = SPMM_THREADS / LPR;
    const int sub = threadIdx.x % LPR;
    const int64_t row = (int64_t)blockIdx.x * GROUPS + threadIdx.x / LPR;
    if (row >= n_out) return; // whole group exits together (LPR divides 32, group-uniform)
    // active-lane mask of this group inside its warp
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x & 31) / LPR * LPR));
    const int nvec = (F + W - 1) / W; // vectors per row

    V acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = T::zero();

    const int32_t begin = ptr[row], end = ptr[row + 1];
    for (int32_t k = begin; k < end; k += LPR) {
        int32_t my_c = 0;
        float my_a = 0.f;
        if (k + sub < end) {
            my_c = ld_stream_i32(idx + k + sub);
            if (USE_VAL) my_a = ld_stream_f32(val + k + sub);
        }
        const int cnt = min(LPR, end - k);
        for (int j = 0; j < cnt; j += SPMM_U) {
            V p[SPMM_U][VEC];
            float a[SPMM_U];
#pragma unroll
            for (int u = 0; u < SPMM_U; u++) {
                const int jj = j + u;                       // group-uniform
                const int src_lane = jj < LPR ? jj : LPR - 1;
                const int32_t c = __shfl_sync(gmask, my_c, src_lane, LPR);
                a[u] = USE_VAL ? __shfl_sync(gmask, my_a, src_lane, LPR) : 1.f;
                const V *prow = reinterpret_cast<const V *>(P + (int64_t)c * ldp);
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    const int vi = sub + v * LPR;
                    if (jj < cnt && vi < nvec) p[u][v] = prow[vi];
                    else p[u][v] = T::zero();
                }
                if (jj >= cnt) a[u] = 0.f;
            }
#pragma unroll
            for (int u = 0; u < SPMM_U; u++) {
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    if (USE_VAL) T::fma(acc[v], a[u], p[u][v]);
                    else T::add(acc[v], p[u][v]);
                }
            }
        }
    }

    const float *mrow = mask ? mask + row * ldm : nullptr;
    float *yrow = Y + row * ldy;
#pragma unroll
    for (int v = 0; v < VEC; v++) {
        const int vi = sub + v * LPR;
        if (vi >= nvec) continue;
        if constexpr (W == 4) {
            float4 o = acc[v];
            const int c0 = vi * 4;
            o.x = epilogue1(o.x, c0 + 0, F, bias, relu, mrow);
            o.y = epilogue1(o.y, c0 + 1, F, bias, relu, mrow);
            o.z = epilogue1(o.z, c0 + 2, F, bias, relu, mrow);
            o.w = epilogue1(o.w, c0 + 3, F, bias, relu, mrow);
            reinterpret_cast<float4 *>(yrow)[vi] = o;
        } else {
            yrow[vi] = epilogue1(acc[v], vi, F, bias, relu, mrow);
        }
    }
}

template <typename V, int LPR, int VEC>
static int launch_rows(gnn_ctx *ctx, int32_t n_out, const int32_t *ptr, const int32_t *idx, const float *val,
                       const float *P, int64_t ldp, int32_t F, float *Y, int64_t ldy, const float *bias, int relu,
                       const float *mask, int64_t ldm) {
    constexpr int GROUPS = SPMM_THREADS / LPR;
    const unsigned grid = (unsigned)ceil_div(n_out, GROUPS);
    if (val)
        spmm_rows_kernel<V, LPR, VEC, true><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n_out, ptr, idx, val, P, ldp, F, Y,
                                                                                   ldy, bias, relu, mask, ldm);
    else
        spmm_rows_kernel<V, LPR, VEC, false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n_out, ptr, idx, val, P, ldp, F,
                                                                                    Y, ldy, bias, relu, mask, ldm);
    GNN_LAUNCHED(ctx);
    return 0;
}

// Dispatch on width.  Wide rows are processed in column blocks (separate launches on shifted pointers).
int spmm_launch(gnn_ctx *ctx, int32_t n_out, const int32_t *ptr, const int32_t *idx, const float *val,
                int32_t max_nnz_row, const float *P, int64_t ldp, int32_t F, float *Y, int64_t ldy, const float *bias,
                int relu, const float *mask, int64_t ldm) {
    (void)max_nnz_row;
    if (n_out <= 0 || F <= 0) return 0;
    const bool vec_ok = ((uintptr_t)P % 16 == 0) && ((uintptr_t)Y % 16 == 0) && (ldp % 4 == 0) && (ldy % 4 == 0) &&
                        ldp >= round_up(F, 4) && ldy >= round_up(F, 4);
    const int32_t block_cols = vec_ok ? 512 : 256;
    for (int32_t c0 = 0; c0 < F; c0 += block_cols) {
        const int32_t f = F - c0 < block_cols ? F - c0 : block_cols;
        const float *Pc = P + c0;
        float *Yc = Y + c0;
        const float *bc = bias ? bias + c0 : nullptr;
        const float *mc = mask ? mask + c0 : nullptr;
#define GO(V, LPR, VEC) GNN_TRY((launch_rows<V, LPR, VEC>(ctx, n_out, ptr, idx, val, Pc, ldp, f, Yc, ldy, bc, relu, mc, ldm)))
        if (vec_ok) {
            const int nv = (f + 3) / 4;
            if (nv <= 4) GO(float4, 4, 1);
            else if (nv <= 8) GO(float4, 8, 1);
            else if (nv <= 16) GO(float4, 16, 1);
            else if (nv <= 32) GO(float4, 32, 1);
            else if (nv <= 64) GO(float4, 32, 2);
            else GO(float4, 32, 4);
        } else {
            if (f <= 8) GO(float, 8, 1);
            else if (f <= 16) GO(float, 16, 1);
            else if (f <= 32) GO(float, 32, 1);
            else if (f <= 64) GO(float, 32, 2);
            else if (f <= 128) GO(float, 32, 4);
            else GO(float, 32, 8);
        }
#undef GO
    }
    return 0;
}

} // namespace gnn

using namespace gnn;

extern "C" {

int gnn_set_spmm_variant(gnn_ctx_t *ctx, int variant) {
    GNN_REQUIRE(ctx && variant >= 0 && variant <= 2, "gnn_set_spmm_variant: bad argument");
    ctx->spmm_variant = variant;
    return 0;
}

int gnn_spmm_fwd(gnn_ctx_t *ctx, const gnn_graph_t *g, const float *P, int64_t ldp, int32_t F, float *Y, int64_t ldy,
                 const float *bias, int relu, const float *mask, int64_t ldm, int use_values) {
    GNN_REQUIRE(ctx && g && P && Y, "gnn_spmm_fwd: NULL argument");
    GNN_REQUIRE(F > 0 && ldp >= F && ldy >= F, "tensors are not compatible, tensors should of shape [...,A,B] and [...,B,A]");
    GNN_REQUIRE(!use_values || g->val, "gnn_spmm_fwd: edge values not built (call gnn_graph_normalize)");
    GNN_REQUIRE(P != Y, "gnn_spmm_fwd: in-place aggregation is not supported");
    return spmm_launch(ctx, g->n_rows, g->rowptr, g->colidx, use_values ? g->val : nullptr, g->max_row_nnz, P, ldp, F,
                       Y, ldy, bias, relu, mask, ldm);
}

int gnn_spmm_bwd(gnn_ctx_t *ctx, const gnn_graph_t *g, const float *dZ, int64_t ldz, int32_t F, float *dP, int64_t ldp,
                 const float *mask, int64_t ldm, int use_values) {
    GNN_REQUIRE(ctx && g && dZ && dP, "gnn_spmm_bwd: NULL argument");
    GNN_REQUIRE(F > 0 && ldz >= F && ldp >= F, "tensors are not compatible, tensors should of shape [...,A,B] and [...,B,A]");
    GNN_REQUIRE(g->colptr, "gnn_spmm_bwd: CSC not built (call gnn_graph_build_csc)");
    GNN_REQUIRE(dZ != dP, "gnn_spmm_bwd: in-place aggregation is not supported");
    const bool alias = g->symmetric; // A_hat^T == A_hat structurally and in value: reuse the CSR arrays
    const float *v = nullptr;
    if (use_values) {
        v = alias ? g->val : g->valT;
        GNN_REQUIRE(v, "gnn_spmm_bwd: edge values not built (call gnn_graph_normalize after gnn_graph_build_csc)");
    }
    return spmm_launch(ctx, g->t_rows, alias ? g->rowptr : g->colptr, alias ? g->colidx : g->rowidx, v, g->max_col_nnz,
                       dZ, ldz, F, dP, ldp, nullptr, 0, mask, ldm);
}

} // extern "C"
