// spmm.cu — aggregation kernels K4/K5:  Y = A_hat * P  over a CSR (forward) or over the CSC of A_hat
// (= CSR of A_hat^T, backward), with the bias / ReLU / ReLU-mask epilogue of K7 fused in.
//
// HBM/L2-gather bound.  Design (B200):
//   - LPR lanes hold one F-wide row as VEC vectors (128-bit when the leading dimensions allow) of accumulators
//     in registers, so each output row is written once and there are NO atomics: deterministic;
//   - the lanes load a batch of (column, value) pairs at a time, coalesced, and broadcast them with shuffles;
//     feature rows are gathered with U independent 128-bit loads in flight per lane;
//   - index/value streams use the no-allocate path (read once), feature rows the default path (L2 reuse);
//   - two work distributions: equal nonzero chunks per WARP with a fixed-order fix-up of the rows cut by chunk
//     boundaries (default: balanced whatever the degree skew, hub rows spread over warps), or one row per lane
//     group (matrices with empty rows, ablation).
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

// output row pointer (see YDest in common.cuh)
__device__ __forceinline__ float *yd_row(const YDest &d, int32_t row, int64_t ldy) {
    if (d.row_map) row = __ldg(d.row_map + row);
    if (d.rows_per == 0) return d.base[0] + (int64_t)row * ldy;
    const int32_t q = row / d.rows_per;
    return d.base[q] + (int64_t)(row - q * d.rows_per) * ldy;
}

constexpr int SPMM_THREADS = 256;
// Loads in flight per lane (U nonzeros x VEC vectors) and whether the next (column, value) batch is requested
// before the current one is consumed.  Measured on B200 (tools/spmm_probe.py, products-shaped, merge kernel):
// 16 B/lane/nonzero: U=4 + prefetch best (F=100: 5.31 ms vs 6.05 without prefetch, 6.32 with U=8);
// 32 B/lane/nonzero (F=256): U=4 without prefetch best (10.6-10.8 ms vs 11.3 with, 13.4 with U=8).
template <typename V, int VEC> struct Tune {
    static constexpr int BYTES = (int)sizeof(V) * VEC;
    static constexpr int U = BYTES >= 64 ? 2 : 4;
    static constexpr bool PF = BYTES <= 16;
};

// acc += sum over k in [begin, end) of val[k] * P[idx[k], :]   for the lane group this thread belongs to.
// A group is S slots of LPR lanes (G = S*LPR lanes, `lig` = lane index inside the group): the LPR lanes of a slot
// cover one feature row, and the S slots work on S consecutive nonzeros at once — narrow rows (F <= 64) keep the
// whole warp on one row range that way instead of running two or more divergent groups per warp.
// The group's lanes fetch G (column, value) pairs at a time (coalesced, streaming; with PF the NEXT batch is
// requested before the current one is consumed) and broadcast them with shuffles; every lane keeps U x VEC 128-bit
// gathers in flight.  A slot accumulates its nonzeros in stored (ascending k) order; the caller combines the slots
// with reduce_slots (fixed tree): deterministic.
template <typename V, int LPR, int S, int VEC, bool USE_VAL, int U, bool PF>
__device__ __forceinline__ void accumulate_range(V (&acc)[VEC], const int32_t *__restrict__ idx,
                                                 const float *__restrict__ val, int32_t begin, int32_t end,
                                                 const float *__restrict__ P, int64_t ldp, int nvec, int lig,
                                                 unsigned gmask) {
    using T = VecTraits<V>;
    constexpr int G = LPR * S;
    const int sub = lig % LPR, slot = lig / LPR;
    int32_t nx_c = 0;
    float nx_a = 0.f;
    if (PF && begin + lig < end) {
        nx_c = ld_stream_i32(idx + begin + lig);
        if (USE_VAL) nx_a = ld_stream_f32(val + begin + lig);
    }
    for (int32_t k = begin; k < end; k += G) {
        int32_t my_c = nx_c;
        float my_a = nx_a;
        if (PF) {
            if (k + G + lig < end) {
                nx_c = ld_stream_i32(idx + k + G + lig);
                if (USE_VAL) nx_a = ld_stream_f32(val + k + G + lig);
            }
        } else if (k + lig < end) {
            my_c = ld_stream_i32(idx + k + lig);
            if (USE_VAL) my_a = ld_stream_f32(val + k + lig);
        }
        const int cnt = min(G, end - k);
        for (int j = 0; j < cnt; j += S * U) {
            V p[U][VEC];
            float a[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int jj = j + u * S + slot;            // uniform inside a slot
                const int src_lane = jj < G ? jj : G - 1;
                const int32_t c = __shfl_sync(gmask, my_c, src_lane, G);
                a[u] = USE_VAL ? __shfl_sync(gmask, my_a, src_lane, G) : 1.f;
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
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int v = 0; v < VEC; v++) {
                    if (USE_VAL) T::fma(acc[v], a[u], p[u][v]);
                    else T::add(acc[v], p[u][v]);
                }
            }
        }
    }
}

// sum the S slots of a group (butterfly over lane offsets LPR, 2 LPR, ...): every lane ends with the total
template <typename V, int LPR, int S, int VEC>
__device__ __forceinline__ void reduce_slots(V (&acc)[VEC], unsigned gmask) {
    if constexpr (S > 1) {
#pragma unroll
        for (int off = LPR; off < LPR * S; off <<= 1) {
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                if constexpr (VecTraits<V>::W == 4) {
                    acc[v].x += __shfl_xor_sync(gmask, acc[v].x, off, LPR * S);
                    acc[v].y += __shfl_xor_sync(gmask, acc[v].y, off, LPR * S);
                    acc[v].z += __shfl_xor_sync(gmask, acc[v].z, off, LPR * S);
                    acc[v].w += __shfl_xor_sync(gmask, acc[v].w, off, LPR * S);
                } else {
                    acc[v] += __shfl_xor_sync(gmask, acc[v], off, LPR * S);
                }
            }
        }
    }
}

template <typename V, int LPR, int VEC>
__device__ __forceinline__ void store_row(const V (&acc)[VEC], float *__restrict__ yrow, int nvec, int sub, int32_t F,
                                          const float *__restrict__ bias, int relu, const float *__restrict__ mrow) {
    constexpr int W = VecTraits<V>::W;
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

// ---- variant 1: one output row per lane group ------------------------------------------------------------------
// V = float4: requires P/Y 16-byte aligned and ldp/ldy multiples of 4 (padding columns may be touched).
// V = float : no alignment requirement.
// MULTI = false: the plain output matrix Y (the single-GPU / row-partition kernels, exactly the r1 code: routing the
// plain case through the destination table as well cost 9 % on the F = 256 launch); MULTI = true: rows go to YDest.
template <typename V, int LPR, int VEC, bool USE_VAL, int U, bool PF, bool MULTI>
__global__ void __launch_bounds__(SPMM_THREADS)
    spmm_rows_kernel(int32_t n_out, const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx,
                     const float *__restrict__ val, const float *__restrict__ P, int64_t ldp, int32_t F,
                     float *__restrict__ Y, const __grid_constant__ YDest yd, int64_t ldy, const float *__restrict__ bias, int relu,
                     const float *__restrict__ mask, int64_t ldm) {
    using T = VecTraits<V>;
    constexpr int GROUPS = SPMM_THREADS / LPR;
    const int sub = threadIdx.x % LPR;
    const int64_t row = (int64_t)blockIdx.x * GROUPS + threadIdx.x / LPR;
    if (row >= n_out) return; // whole group exits together (LPR divides 32, group-uniform)
    // active-lane mask of this group inside its warp
    const unsigned gmask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << ((threadIdx.x & 31) / LPR * LPR));
    const int nvec = (F + T::W - 1) / T::W; // vectors per row

    V acc[VEC];
#pragma unroll
    for (int v = 0; v < VEC; v++) acc[v] = T::zero();
    accumulate_range<V, LPR, 1, VEC, USE_VAL, U, PF>(acc, idx, val, ptr[row], ptr[row + 1], P, ldp, nvec, sub, gmask);
    float *yrow;
    if constexpr (MULTI) yrow = yd_row(yd, (int32_t)row, ldy);
    else yrow = Y + row * ldy;
    store_row<V, LPR, VEC>(acc, yrow, nvec, sub, F, bias, relu, mask ? mask + row * ldm : nullptr);
}

// ---- variant 2: nonzero-balanced (merge-path style) ------------------------------------------------------------
// The nonzero range [0, nnz) is cut into chunks of MERGE_CHUNK; a lane group owns one chunk and walks the rows that
// intersect it (start row by binary search on ptr).  Rows entirely inside the chunk are finished in place.  A row
// cut by a chunk boundary leaves partial sums in a workspace: `tail[g]` when the row continues past chunk g,
// `head[g]` when it started before chunk g and ends inside it.  spmm_merge_fixup_kernel then adds, for every head,
// the tails of the preceding chunks of the same row in ascending chunk order, applies the epilogue and writes the
// row — fixed order, no atomics.  Hub rows are thus spread over as many groups as they have chunks.
// Requires every row to hold at least one nonzero (the dispatcher checks min_nnz_row >= 1).
constexpr int MERGE_CHUNK_MIN = 256;

// Resident CTAs per SM the register allocation must allow.  The narrow configurations (one 128-bit vector per lane)
// are latency-bound gathers: ncu on the F=16 launch showed 64 registers -> 4 CTAs -> 46 % of the warp slots active
// with DRAM at 50 % and every pipe under 50 %, so they trade registers for resident warps.
// Measured on one B200 (tools/spmm_width_probe.py, products-shaped, ms per launch; profiles/r2_spmm_occupancy_ab.md):
//   width            12     16     24     32     48     64    100    128
//   4 CTAs (r1)    1.42   1.46   2.23   2.15   3.41   3.35   5.27   5.29
//   5 CTAs         1.32   1.34   1.97   1.88   2.99   2.96   5.28   5.28
//   6 CTAs         1.54   1.53   2.10   1.91   3.05   3.04   5.41   5.45   (40 registers: ~100 bytes of spills)
// The wider configurations (two or four vectors per lane) are HBM-bound already and keep the default allocation.
#ifndef SPMM_MIN_CTAS_NARROW
#define SPMM_MIN_CTAS_NARROW 5
#endif
template <typename V, int VEC> struct Occupancy {
    static constexpr int MIN_CTAS = (sizeof(V) * VEC <= 16) ? SPMM_MIN_CTAS_NARROW : 0; // 0 = unspecified
};

template <typename V, int LPR, int VEC, bool USE_VAL, int U, bool PF, bool MULTI>
__global__ void __launch_bounds__(SPMM_THREADS, (Occupancy<V, VEC>::MIN_CTAS))
    spmm_merge_kernel(int32_t n_out, int32_t k_base, int32_t nnz, int32_t n_chunks, int32_t MERGE_CHUNK,
                      const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx, const float *__restrict__ val,
                      const float *__restrict__ P, int64_t ldp, int32_t F, float *__restrict__ Y,
                      const __grid_constant__ YDest yd, int64_t ldy,
                      const float *__restrict__ bias, int relu, const float *__restrict__ mask, int64_t ldm,
                      float *__restrict__ head, float *__restrict__ tail, int32_t *__restrict__ head_row,
                      int32_t *__restrict__ tail_row, int32_t ldw) {
    using T = VecTraits<V>;
    constexpr int S = 32 / LPR;                  // the whole warp owns one chunk: S nonzeros in flight side by side
    constexpr int GROUPS = SPMM_THREADS / 32;
    const int lig = threadIdx.x & 31, sub = lig % LPR, slot = lig / LPR;
    const int32_t g = blockIdx.x * GROUPS + (threadIdx.x >> 5);
    if (g >= n_chunks) return;
    const unsigned gmask = 0xffffffffu;
    const int nvec = (F + T::W - 1) / T::W;
    const int nvec_st = slot == 0 ? nvec : 0;    // after reduce_slots every slot holds the sum; slot 0 stores it
    // nonzeros [k_base, nnz) belong to rows [0, n_out) of `ptr` (a row range of a larger matrix keeps absolute offsets)
    const int32_t k0 = k_base + g * MERGE_CHUNK, k1 = min(nnz, k0 + MERGE_CHUNK);

    // last row r with ptr[r] <= k0 (rows are non-empty, so it is the row that holds nonzero k0)
    int32_t lo = 0, hi = n_out; // invariant: ptr[lo] <= k0 < ptr[hi]
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) <= k0) lo = mid;
        else hi = mid;
    }
    int32_t row = lo, k = k0;
    int32_t hrow = -1, trow = -1;
    int32_t rb = __ldg(ptr + row), re = __ldg(ptr + row + 1);
    while (k < k1) {
        const int32_t seg_end = min(re, k1);
        const int32_t re_next = (seg_end == re && row + 2 <= n_out) ? __ldg(ptr + row + 2) : re; // prefetch
        V acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; v++) acc[v] = T::zero();
        accumulate_range<V, LPR, S, VEC, USE_VAL, U, PF>(acc, idx, val, k, seg_end, P, ldp, nvec, lig, gmask);
        reduce_slots<V, LPR, S, VEC>(acc, gmask);
        const bool starts = (k == rb), ends = (seg_end == re);
        if (starts && ends) {
            float *yrow;
            if constexpr (MULTI) yrow = yd_row(yd, row, ldy);
            else yrow = Y + (int64_t)row * ldy;
            store_row<V, LPR, VEC>(acc, yrow, nvec_st, sub, F, bias, relu,
                                   mask ? mask + (int64_t)row * ldm : nullptr);
        } else {
            float *dst = (ends ? head : tail) + (int64_t)g * ldw;
            if (ends) hrow = row;
            else trow = row;
#pragma unroll
            for (int v = 0; v < VEC; v++) {
                const int vi = sub + v * LPR;
                if (vi < nvec_st) reinterpret_cast<V *>(dst)[vi] = acc[v];
            }
        }
        k = seg_end;
        row++;
        rb = re;
        re = re_next;
    }
    if (lig == 0) {
        head_row[g] = hrow;
        tail_row[g] = trow;
    }
}

// one warp per chunk that finishes a cut row
__global__ void __launch_bounds__(256)
    spmm_merge_fixup_kernel(int32_t n_chunks, int32_t F, const float *__restrict__ head, const float *__restrict__ tail,
                            const int32_t *__restrict__ head_row, const int32_t *__restrict__ tail_row, int32_t ldw,
                            const __grid_constant__ YDest yd, int64_t ldy, const float *__restrict__ bias, int relu,
                            const float *__restrict__ mask, int64_t ldm) {
    const int32_t g = (int32_t)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (g >= n_chunks) return;
    const int32_t row = head_row[g];
    if (row < 0) return;
    int32_t first = g;
    while (first > 0 && tail_row[first - 1] == row) first--;
    const float *mrow = mask ? mask + (int64_t)row * ldm : nullptr;
    float *yrow = yd_row(yd, row, ldy);
    for (int32_t c = lane; c < F; c += 32) {
        float s = 0.f;
        for (int32_t t = first; t < g; t++) s += tail[(int64_t)t * ldw + c];
        s += head[(int64_t)g * ldw + c];
        yrow[c] = epilogue1(s, c, F, bias, relu, mrow);
    }
}

// ---- variant 2a: the nonzero-balanced walk with the gathers staged through shared memory ----------------------------
// The one-vector-per-lane widths (F <= 128) are latency-bound gathers: a lane can only keep U = 4 feature-row loads in
// flight in registers, and every row segment restarts the chain "load (column, value) -> load the rows -> accumulate".
// Here every lane owns a FIFO of D 16-byte slots in shared memory that it fills with cp.async (LDGSTS: no destination
// registers, so the depth costs nothing but shared memory) and drains D steps later with a 128-bit shared load.  The
// fill side ("producer") walks the chunk's nonzeros D steps ahead of the accumulate side ("consumer") ACROSS row
// boundaries, and streams the (column, value) pairs in chunk order, so the gather pipeline never drains at a row end.
// Both sides are the same warp and use the same row-aligned step sequence (step = S consecutive nonzeros of one row,
// one per slot), hence the summation order — and every bit of the result — is that of spmm_merge_kernel.
// A lane only ever reads back what it copied itself: no barrier of any kind is needed.
//
// MEASURED (tools/spmm_async_probe.py, products-shaped, 1x B200, ms per launch; profiles/r2b_spmm_async.md): bit-identical
// at every width, but only the narrowest width gains —
//   width               16     32     47     64    100    128
//   register gathers   1.36   1.87   2.99   2.95   5.27   5.27
//   staged, depth 8    1.26   1.87   3.13   3.09   5.59   5.83      (depth 12, 3 resident CTAs: 1.58 ... 6.29)
// so doubling the loads in flight does not buy bandwidth: the narrow aggregations are not short of memory-level
// parallelism (a 192-byte row costs what a 256-byte row costs — the limit is per request, not per byte), and the staged
// walk pays ~2.7x the instructions per nonzero plus the shared-memory round trip.  NOT the default: kept behind
// GNN_SPMM_ASYNC=8 (context option) as the recorded shared-memory-staging ablation the north star asks for.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <int LPR, int D, bool USE_VAL, bool MULTI>
__global__ void __launch_bounds__(SPMM_THREADS, SPMM_MIN_CTAS_NARROW)
    spmm_merge_async_kernel(int32_t n_out, int32_t k_base, int32_t nnz, int32_t n_chunks, int32_t MERGE_CHUNK,
                            const int32_t *__restrict__ ptr, const int32_t *__restrict__ idx, const float *__restrict__ val,
                            const float *__restrict__ P, int64_t ldp, int32_t F, float *__restrict__ Y,
                            const __grid_constant__ YDest yd, int64_t ldy, const float *__restrict__ bias, int relu,
                            const float *__restrict__ mask, int64_t ldm, float *__restrict__ head, float *__restrict__ tail,
                            int32_t *__restrict__ head_row, int32_t *__restrict__ tail_row, int32_t ldw) {
    using V = float4;
    using T = VecTraits<V>;
    constexpr int S = 32 / LPR;
    constexpr int GROUPS = SPMM_THREADS / 32;
    extern __shared__ __align__(16) uint8_t fifo_raw[];
    float4 *fifo_p = reinterpret_cast<float4 *>(fifo_raw);                         // [D][SPMM_THREADS]
    float *fifo_a = reinterpret_cast<float *>(fifo_raw + (size_t)D * SPMM_THREADS * 16); // [D][SPMM_THREADS]
    const int lig = threadIdx.x & 31, sub = lig % LPR, slot = lig / LPR;
    const int32_t g = blockIdx.x * GROUPS + (threadIdx.x >> 5);
    if (g >= n_chunks) return;
    const unsigned gmask = 0xffffffffu;
    const int nvec = (F + 3) / 4;
    const int nvec_st = slot == 0 ? nvec : 0;
    const bool lane_ok = sub < nvec;
    const int32_t k0 = k_base + g * MERGE_CHUNK, k1 = min(nnz, k0 + MERGE_CHUNK);
    const uint32_t my_p = (uint32_t)__cvta_generic_to_shared(fifo_p + threadIdx.x);

    int32_t lo = 0, hi = n_out; // invariant: ptr[lo] <= k0 < ptr[hi]
    while (hi - lo > 1) {
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(ptr + mid) <= k0) lo = mid;
        else hi = mid;
    }
    // ---- producer cursor: next step = nonzeros p_k + slot of row p_row, while < p_end (the row's end inside the chunk)
    int32_t p_row = lo, p_k = k0;
    int32_t p_end = min(__ldg(ptr + p_row + 1), k1);
    // (column, value) pairs in chunk order: cur covers [bk, bk + 32), nxt the 32 after it
    int32_t bk = k0;
    int32_t cur_c = 0, nxt_c = 0;
    float cur_a = 1.f, nxt_a = 1.f;
    if (bk + lig < nnz) {
        cur_c = ld_stream_i32(idx + bk + lig);
        if (USE_VAL) cur_a = ld_stream_f32(val + bk + lig);
    }
    if (bk + 32 + lig < nnz) {
        nxt_c = ld_stream_i32(idx + bk + 32 + lig);
        if (USE_VAL) nxt_a = ld_stream_f32(val + bk + 32 + lig);
    }
    auto produce = [&](int fslot) { // one step into FIFO slot fslot — or an empty group once the chunk is exhausted
        if (p_k < k1) {
            const int32_t kk = p_k + slot;
            const bool live = kk < p_end;
            const int rel = (kk - bk) & 63; // < 40 for live lanes
            const int32_t c0 = __shfl_sync(gmask, cur_c, rel & 31), c1 = __shfl_sync(gmask, nxt_c, rel & 31);
            const int32_t c = rel < 32 ? c0 : c1;
            float a = 1.f;
            if (USE_VAL) {
                const float a0 = __shfl_sync(gmask, cur_a, rel & 31), a1 = __shfl_sync(gmask, nxt_a, rel & 31);
                a = rel < 32 ? a0 : a1;
            }
            const bool fetch = live && lane_ok;
            const float *src = fetch ? P + (int64_t)c * ldp + sub * 4 : P;
            cp_async16(my_p + (uint32_t)fslot * (SPMM_THREADS * 16), src, fetch ? 16u : 0u); // 0 bytes: the slot is zero-filled
            fifo_a[fslot * SPMM_THREADS + threadIdx.x] = live ? a : 0.f;
            p_k += S;
            if (p_k >= p_end) { // row finished (inside this chunk): the next step starts the next row
                p_k = p_end;
                if (p_k < k1) {
                    p_row++;
                    if ((p_row & 31) == 0) prefetch_l1(ptr + min(p_row + 32, n_out)); // next line of row pointers
                    p_end = min(__ldg(ptr + p_row + 1), k1);
                }
            }
            if (p_k - bk >= 32) { // slide the (column, value) window; the batch after next is requested now
                bk += 32;
                cur_c = nxt_c;
                cur_a = nxt_a;
                if (bk + 32 + lig < nnz) {
                    nxt_c = ld_stream_i32(idx + bk + 32 + lig);
                    if (USE_VAL) nxt_a = ld_stream_f32(val + bk + 32 + lig);
                }
            }
        }
        cp_async_commit();
    };

    // ---- consumer cursor (same step sequence, D - 1 steps behind)
    int32_t row = lo, k = k0;
    int32_t re = __ldg(ptr + row + 1);
    int32_t seg_end = min(re, k1);
    bool starts = __ldg(ptr + row) == k0; // only the chunk's first row can have begun in an earlier chunk
    int32_t hrow = -1, trow = -1;
    V acc = T::zero();
#pragma unroll
    for (int i = 0; i < D - 1; i++) produce(i);
    int fs = 0; // FIFO slot of the step consumed next
    while (k < k1) {
        produce(fs == 0 ? D - 1 : fs - 1); // the slot consumed in the previous iteration
        cp_async_wait<D - 1>();
        const float4 pv = fifo_p[fs * SPMM_THREADS + threadIdx.x];
        const float av = fifo_a[fs * SPMM_THREADS + threadIdx.x];
        T::fma(acc, av, pv);
        fs = fs + 1 == D ? 0 : fs + 1;
        k += S;
        if (k >= seg_end) { // the row's nonzeros inside this chunk are summed
            V accv[1] = {acc};
            reduce_slots<V, LPR, S, 1>(accv, gmask);
            acc = accv[0];
            const bool ends = (seg_end == re);
            if (starts && ends) {
                float *yrow;
                if constexpr (MULTI) yrow = yd_row(yd, row, ldy);
                else yrow = Y + (int64_t)row * ldy;
                store_row<V, LPR, 1>(accv, yrow, nvec_st, sub, F, bias, relu, mask ? mask + (int64_t)row * ldm : nullptr);
            } else {
                float *dst = (ends ? head : tail) + (int64_t)g * ldw;
                if (ends) hrow = row;
                else trow = row;
                if (sub < nvec_st) reinterpret_cast<V *>(dst)[sub] = acc;
            }
            acc = T::zero();
            k = seg_end;
            starts = true;
            if (k < k1) {
                row++;
                re = __ldg(ptr + row + 1);
                seg_end = min(re, k1);
            }
        }
    }
    cp_async_wait<0>();
    if (lig == 0) {
        head_row[g] = hrow;
        tail_row[g] = trow;
    }
}

// nonzeros per warp of the merge kernel; a small launch (e.g. a row block of a partition) is cut finer so that it
// still spans >= 4 waves of resident warps
static inline int32_t merge_chunk(const gnn_ctx *ctx, int lpr, int64_t nnz) {
    (void)lpr;
    if (ctx->spmm_chunk > 0) return ctx->spmm_chunk;
    int64_t chunk = 1024;
    const int64_t want_chunks = 4ll * ctx->sm_count * 64;
    if (nnz / chunk < want_chunks) chunk = nnz / want_chunks / 64 * 64;
    return (int32_t)(chunk < 128 ? 128 : chunk);
}

template <typename V, int LPR, int VEC, int U, bool PF>
static int launch_rows(gnn_ctx *ctx, int32_t n_out, const int32_t *ptr, const int32_t *idx, const float *val,
                       const float *P, int64_t ldp, int32_t F, const YDest &Y, int64_t ldy, const float *bias, int relu,
                       const float *mask, int64_t ldm) {
    constexpr int GROUPS = SPMM_THREADS / LPR;
    const unsigned grid = (unsigned)ceil_div(n_out, GROUPS);
    float *Y0 = Y.base[0];
#define ROWS_GO(UV, MU) spmm_rows_kernel<V, LPR, VEC, UV, U, PF, MU><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n_out, ptr, idx, val, P, ldp, F, Y0, Y, ldy, bias, relu, mask, ldm)
    if (Y.rows_per || Y.row_map) { if (val) ROWS_GO(true, true); else ROWS_GO(false, true); }
    else { if (val) ROWS_GO(true, false); else ROWS_GO(false, false); }
#undef ROWS_GO
    GNN_LAUNCHED(ctx);
    return 0;
}

template <typename V, int LPR, int VEC, int U, bool PF>
static int launch_merge(gnn_ctx *ctx, int32_t n_out, int64_t k_base, int64_t nnz, const int32_t *ptr, const int32_t *idx,
                        const float *val, const float *P, int64_t ldp, int32_t F, const YDest &Y, int64_t ldy,
                        const float *bias, int relu, const float *mask, int64_t ldm) {
    constexpr int GROUPS = SPMM_THREADS / 32; // one warp per chunk
    const int32_t MERGE_CHUNK = merge_chunk(ctx, LPR, nnz - k_base);
    const int32_t n_chunks = (int32_t)ceil_div(nnz - k_base, MERGE_CHUNK);
    const int32_t ldw = (int32_t)round_up(F, 4);
    void *ws = nullptr;
    const size_t part = (size_t)n_chunks * ldw * 4;
    GNN_TRY(ctx->workspace(2 * part + (size_t)n_chunks * 8 + 64, &ws));
    float *head = (float *)ws, *tail = head + (size_t)n_chunks * ldw;
    int32_t *head_row = (int32_t *)(tail + (size_t)n_chunks * ldw), *tail_row = head_row + n_chunks;
    const unsigned grid = (unsigned)ceil_div(n_chunks, GROUPS);
    float *Y0 = Y.base[0];
#define MERGE_GO(UV, MU)                                                                                            \
    spmm_merge_kernel<V, LPR, VEC, UV, U, PF, MU><<<grid, SPMM_THREADS, 0, ctx->stream>>>(                          \
        n_out, (int32_t)k_base, (int32_t)nnz, n_chunks, MERGE_CHUNK, ptr, idx, val, P, ldp, F, Y0, Y, ldy, bias, relu, mask, ldm, head, \
        tail, head_row, tail_row, ldw)
    if (Y.rows_per || Y.row_map) { if (val) MERGE_GO(true, true); else MERGE_GO(false, true); }
    else { if (val) MERGE_GO(true, false); else MERGE_GO(false, false); }
#undef MERGE_GO
    GNN_LAUNCHED(ctx);
    spmm_merge_fixup_kernel<<<(unsigned)ceil_div((int64_t)n_chunks * 32, 256), 256, 0, ctx->stream>>>(
        n_chunks, F, head, tail, head_row, tail_row, ldw, Y, ldy, bias, relu, mask, ldm);
    GNN_LAUNCHED(ctx);
    return 0;
}

// the shared-memory-staged walk (spmm_merge_async_kernel) for the one-vector-per-lane widths; D = FIFO depth
template <int LPR, int D>
static int launch_merge_async(gnn_ctx *ctx, int32_t n_out, int64_t k_base, int64_t nnz, const int32_t *ptr, const int32_t *idx,
                              const float *val, const float *P, int64_t ldp, int32_t F, const YDest &Y, int64_t ldy,
                              const float *bias, int relu, const float *mask, int64_t ldm) {
    constexpr int GROUPS = SPMM_THREADS / 32;
    const int32_t MERGE_CHUNK = merge_chunk(ctx, LPR, nnz - k_base);
    const int32_t n_chunks = (int32_t)ceil_div(nnz - k_base, MERGE_CHUNK);
    const int32_t ldw = (int32_t)round_up(F, 4);
    void *ws = nullptr;
    const size_t part = (size_t)n_chunks * ldw * 4;
    GNN_TRY(ctx->workspace(2 * part + (size_t)n_chunks * 8 + 64, &ws));
    float *head = (float *)ws, *tail = head + (size_t)n_chunks * ldw;
    int32_t *head_row = (int32_t *)(tail + (size_t)n_chunks * ldw), *tail_row = head_row + n_chunks;
    const unsigned grid = (unsigned)ceil_div(n_chunks, GROUPS);
    constexpr size_t smem = (size_t)D * SPMM_THREADS * 20;
    float *Y0 = Y.base[0];
#define ASYNC_GO(UV, MU)                                                                                            \
    do {                                                                                                            \
        auto kern = spmm_merge_async_kernel<LPR, D, UV, MU>;                                                        \
        static uint64_t attr_set = 0; /* per device ordinal */                                                      \
        if (!(attr_set >> (ctx->device & 63) & 1)) {                                                                \
            GNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     \
            GNN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));        \
            attr_set |= 1ull << (ctx->device & 63);                                                                 \
        }                                                                                                           \
        kern<<<grid, SPMM_THREADS, smem, ctx->stream>>>(n_out, (int32_t)k_base, (int32_t)nnz, n_chunks, MERGE_CHUNK, ptr, idx, \
                                                        val, P, ldp, F, Y0, Y, ldy, bias, relu, mask, ldm, head, tail,     \
                                                        head_row, tail_row, ldw);                                   \
    } while (0)
    if (Y.rows_per || Y.row_map) { if (val) ASYNC_GO(true, true); else ASYNC_GO(false, true); }
    else { if (val) ASYNC_GO(true, false); else ASYNC_GO(false, false); }
#undef ASYNC_GO
    GNN_LAUNCHED(ctx);
    spmm_merge_fixup_kernel<<<(unsigned)ceil_div((int64_t)n_chunks * 32, 256), 256, 0, ctx->stream>>>(
        n_chunks, F, head, tail, head_row, tail_row, ldw, Y, ldy, bias, relu, mask, ldm);
    GNN_LAUNCHED(ctx);
    return 0;
}

// Variant choice (ctx->spmm_variant: 0 auto, 1 rows, 2 merge) — by DEGREE SKEW = longest row / mean row length.
// One-row-per-group (rows kernel) has no chunk bookkeeping, no partial-row workspace and no fix-up launch, but a lane
// group streams its row sequentially: a row k times the mean length holds its CTA slot k times as long, and a hub row
// is a serial tail.  The nonzero-balanced merge kernel spreads every row over as many warps as it has 1,024-nonzero
// chunks.  Measured on B200, ms per launch (tools/spmm_probe.py, profiles/r2_spmm_variant_by_skew.md):
//   graph (skew)            width   rows    merge
//   arxiv-shaped (3)          40    0.055   0.093      uniform degrees: rows wins at every width
//                            128    0.111   0.144
//                            256    0.242   0.263
//   pubmed-/cora-shaped (<4)  any   0.016   0.043      launch-bound: one launch instead of two
//   products-shaped (260)    100    5.64    5.31       power law: merge wins from the middle widths up
//                            256   12.95   10.4
//   reddit-shaped (120)       41    5.43    2.05       hub rows of 60 K nonzeros: merge 2-2.6x faster
//                            128    9.52    4.90
// Hub rows are therefore not staged through shared memory: cutting them into chunks that go through the ordinary
// register-accumulating path (and summing the per-chunk partials in fixed order) already runs the Reddit-shaped
// aggregation at 1.9x the HBM copy peak on B_alg (features are L2-resident there), with no smem capacity limit on the
// row length.  The rows kernel also serves matrices with empty rows (the merge walk needs every row to own a nonzero).
constexpr int64_t SPMM_SKEW_MERGE = 16; // longest row >= 16 x the mean -> nonzero-balanced kernel
static bool use_merge(const gnn_ctx *ctx, int64_t nnz, int64_t k_end, int32_t n_out, int32_t min_nnz_row,
                      int32_t max_nnz_row, int lpr) {
    if (min_nnz_row < 1 || k_end >= (1ll << 31) || nnz < 4 * (int64_t)merge_chunk(ctx, lpr, nnz)) return false;
    if (ctx->spmm_variant == 1) return false;
    if (ctx->spmm_variant == 2) return true;
    const int64_t mean = nnz / (n_out > 0 ? n_out : 1) + 1;
    return (int64_t)max_nnz_row >= SPMM_SKEW_MERGE * mean;
}

// Dispatch on width.  Wide rows are processed in column blocks (separate launches on shifted pointers).
// Rows [0, n_out) of `ptr` hold the nonzeros [k_base, nnz) of idx/val (k_base = 0 for a whole matrix; a row range of
// a larger matrix passes ptr + r0 with the absolute offsets ptr[r0], ptr[r1]).
int spmm_launch(gnn_ctx *ctx, int32_t n_out, int64_t k_base, int64_t nnz, const int32_t *ptr, const int32_t *idx,
                const float *val, int32_t min_nnz_row, int32_t max_nnz_row, const float *P, int64_t ldp, int32_t F,
                float *Y, int64_t ldy, const float *bias, int relu, const float *mask, int64_t ldm,
                bool may_touch_padding, const YDest *dest) {
    if (n_out <= 0 || F <= 0) return 0;
    if (dest) Y = dest->base[0]; // alignment check below: every destination shares base[0]'s alignment (same arena offset)
    // The 128-bit path reads columns F..round_up(F,4)-1 of P and WRITES the same columns of Y.  The trainer owns its
    // padded buffers; a public-API caller may pass a column slice of a wider matrix, whose neighbouring columns must
    // survive: such calls (F % 4 != 0 with ld wider than the padded row) take the scalar path.
    const bool vec_ok = ((uintptr_t)P % 16 == 0) && ((uintptr_t)Y % 16 == 0) && (ldp % 4 == 0) && (ldy % 4 == 0) &&
                        ldp >= round_up(F, 4) && ldy >= round_up(F, 4) && (may_touch_padding || F % 4 == 0);
    const int32_t block_cols = vec_ok ? 512 : 256;
    for (int32_t c0 = 0; c0 < F; c0 += block_cols) {
        const int32_t f = F - c0 < block_cols ? F - c0 : block_cols;
        const float *Pc = P + c0;
        YDest Yc;
        Yc.rows_per = dest ? dest->rows_per : 0;
        Yc.row_map = dest ? dest->row_map : nullptr;
        for (int q = 0; q < SPMM_MAX_DEST; q++) Yc.base[q] = dest ? (dest->base[q] ? dest->base[q] + c0 : nullptr) : (q == 0 ? Y + c0 : nullptr);
        const float *bc = bias ? bias + c0 : nullptr;
        const float *mc = mask ? mask + c0 : nullptr;
#define GO2(V, LPR, VEC, U, PF)                                                                                      \
    do {                                                                                                             \
        if (use_merge(ctx, nnz - k_base, nnz, n_out, min_nnz_row, max_nnz_row, LPR)) {                                     \
            if (sizeof(V) == 16 && VEC == 1 && ctx->spmm_async > 0)                                                  \
                GNN_TRY((launch_merge_async<LPR, 8>(ctx, n_out, k_base, nnz, ptr, idx, val, Pc, ldp, f, Yc, ldy, bc, relu, mc, ldm))); \
            else                                                                                                     \
                GNN_TRY((launch_merge<V, LPR, VEC, U, PF>(ctx, n_out, k_base, nnz, ptr, idx, val, Pc, ldp, f, Yc, ldy, bc, relu, mc, ldm))); \
        } else                                                                                                       \
            GNN_TRY((launch_rows<V, LPR, VEC, U, PF>(ctx, n_out, ptr, idx, val, Pc, ldp, f, Yc, ldy, bc, relu, mc, ldm)));  \
    } while (0)
#ifdef GNN_SPMM_TUNE /* experiment build: (U, prefetch) selectable at run time for the float4 kernels */
#define GO(V, LPR, VEC)                                                                                              \
    do {                                                                                                             \
        const int tu = ctx->spmm_tune_u, tp = ctx->spmm_tune_pf;                                                     \
        if (sizeof(V) == 16 && tu == 2 && tp == 0) GO2(V, LPR, VEC, 2, false);                                       \
        else if (sizeof(V) == 16 && tu == 2 && tp == 1) GO2(V, LPR, VEC, 2, true);                                   \
        else if (sizeof(V) == 16 && tu == 4 && tp == 0) GO2(V, LPR, VEC, 4, false);                                  \
        else if (sizeof(V) == 16 && tu == 4 && tp == 1) GO2(V, LPR, VEC, 4, true);                                   \
        else if (sizeof(V) == 16 && tu == 8 && tp == 0) GO2(V, LPR, VEC, 8, false);                                  \
        else if (sizeof(V) == 16 && tu == 8 && tp == 1) GO2(V, LPR, VEC, 8, true);                                   \
        else GO2(V, LPR, VEC, (Tune<V, VEC>::U), (Tune<V, VEC>::PF));                                                               \
    } while (0)
#else
#define GO(V, LPR, VEC) GO2(V, LPR, VEC, (Tune<V, VEC>::U), (Tune<V, VEC>::PF))
#endif
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
#undef GO2
    }
    return 0;
}

// Aggregation restricted to output rows [r0, r1) (the trainer's row blocks).  transpose = 0: A_hat (CSR),
// 1: A_hat^T (CSC, or the CSR arrays again for a symmetric graph).  k0/k1 = ptr[r0]/ptr[r1] (host copies).
// Y and mask point at row r0.
int spmm_rows_range(gnn_ctx *ctx, const gnn_graph *g, int transpose, int32_t r0, int32_t r1, int64_t k0, int64_t k1,
                    const float *P, int64_t ldp, int32_t F, float *Y, int64_t ldy, const float *bias, int relu,
                    const float *mask, int64_t ldm, const YDest *dest) {
    const bool alias = transpose && g->symmetric;
    const int32_t *ptr = (!transpose || alias) ? g->rowptr : g->colptr;
    const int32_t *idx = (!transpose || alias) ? g->colidx : g->rowidx;
    const float *val = !transpose ? g->val : (alias ? (g->valT ? g->valT : g->val) : g->valT);
    const int32_t mn = (!transpose || alias) ? g->min_row_nnz : g->min_col_nnz;
    const int32_t mx = (!transpose || alias) ? g->max_row_nnz : g->max_col_nnz;
    return spmm_launch(ctx, r1 - r0, k0, k1, ptr + r0, idx, val, mn, mx, P, ldp, F, Y, ldy, bias, relu, mask, ldm, true, dest);
}

} // namespace gnn

using namespace gnn;

extern "C" {

int gnn_set_spmm_variant(gnn_ctx_t *ctx, int variant) {
    // variant % 10: 0 auto / 1 rows / 2 merge.  Tuning digits (experiments): (variant / 10) % 10 = loads in flight
    // U (only in GNN_SPMM_TUNE builds), (variant / 100) % 10 = 1 disables the index prefetch (same builds),
    // variant / 1000 = merge chunk in units of 256 nonzeros (0 keeps the default).
    GNN_REQUIRE(ctx && variant >= 0 && variant % 10 <= 2, "gnn_set_spmm_variant: bad argument");
    ctx->spmm_variant = variant % 10;
    ctx->spmm_tune_u = (variant / 10) % 10;
    ctx->spmm_tune_pf = ((variant / 100) % 10) ? 0 : 1;
    const int ch = variant / 1000;
    ctx->spmm_chunk = ch > 0 ? ch * MERGE_CHUNK_MIN : 0;
    return 0;
}

int gnn_graph_spmm_variant(gnn_ctx_t *ctx, const gnn_graph_t *g, int transpose) {
    if (!ctx || !g) return 0;
    const bool alias = transpose && g->symmetric;
    const bool fwd = !transpose || alias;
    const int64_t nnz = fwd ? g->nnz : g->nnz_t;
    return use_merge(ctx, nnz, nnz, fwd ? g->n_rows : g->t_rows, fwd ? g->min_row_nnz : g->min_col_nnz,
                     fwd ? g->max_row_nnz : g->max_col_nnz, 32) ? 2 : 1;
}

int gnn_spmm_fwd(gnn_ctx_t *ctx, const gnn_graph_t *g, const float *P, int64_t ldp, int32_t F, float *Y, int64_t ldy,
                 const float *bias, int relu, const float *mask, int64_t ldm, int use_values) {
    GNN_REQUIRE(ctx && g && P && Y, "gnn_spmm_fwd: NULL argument");
    GNN_REQUIRE(F > 0 && ldp >= F && ldy >= F, "tensors are not compatible, tensors should of shape [...,A,B] and [...,B,A]");
    GNN_REQUIRE(!use_values || g->val, "gnn_spmm_fwd: edge values not built (call gnn_graph_normalize)");
    GNN_REQUIRE(P != Y, "gnn_spmm_fwd: in-place aggregation is not supported");
    const bool pad_ok = ldy == round_up(F, 4) && ldp == round_up(F, 4); // rows are exactly the padded width: padding is the caller's own
    return spmm_launch(ctx, g->n_rows, 0, g->nnz, g->rowptr, g->colidx, use_values ? g->val : nullptr, g->min_row_nnz,
                       g->max_row_nnz, P, ldp, F, Y, ldy, bias, relu, mask, ldm, pad_ok);
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
        v = alias ? (g->valT ? g->valT : g->val) : g->valT; // as-written normalisation keeps a transposed value array
        GNN_REQUIRE(v, "gnn_spmm_bwd: edge values not built (call gnn_graph_normalize after gnn_graph_build_csc)");
    }
    return spmm_launch(ctx, g->t_rows, 0, alias ? g->nnz : g->nnz_t, alias ? g->rowptr : g->colptr,
                       alias ? g->colidx : g->rowidx, v, alias ? g->min_row_nnz : g->min_col_nnz,
                       alias ? g->max_row_nnz : g->max_col_nnz, dZ, ldz, F, dP, ldp, nullptr, 0, mask, ldm,
                       ldz == round_up(F, 4) && ldp == round_up(F, 4));
}

} // extern "C"
