// gemm.cu — dense feature transforms K6 (FP32 FMA path):
//   NT: C[M,N]   = A[M,K]  * B[N,K]^T (+bias)(relu)    forward  P = H W^T            (nn.cpp:205-211)
//   NN: C[M,N]   = A[M,K]  * B[K,N]   (mask)           backward dH = dP W            (operation.h:516-523)
//   TN: C[K1,K2] = A[M,K1]^T * B[M,K2]                 backward dW = dP^T H          (operation.h:524-531,416-433)
// One tiled kernel parameterised on operand layouts; the TN product (reduction over the node dimension,
// millions long) is split over the grid and combined by a fixed-order second pass: no atomics, the result is
// deterministic.  precision=1 (3xTF32 on tcgen05) lives in gemm_tc.cu and falls back here when a shape does not
// fit its tiles.
#include "common.cuh"

namespace gnn {

constexpr int GEMM_THREADS = 256;
constexpr int BM = 128;
constexpr int BK = 16;

struct GemmEpilogue {
    const float *bias; // per output column, or NULL
    int relu;
    const float *mask; // [M, N] with ldm, or NULL: out = mask > 0 ? out : 0
    int64_t ldm;
};

// element (m,k) of op(A) and (k,n) of op(B)
template <bool TA> __device__ __forceinline__ float ldA(const float *A, int64_t lda, int64_t m, int64_t k) {
    return TA ? A[k * lda + m] : A[m * lda + k];
}
template <bool TB> __device__ __forceinline__ float ldB(const float *B, int64_t ldb, int64_t k, int64_t n) {
    return TB ? B[n * ldb + k] : B[k * ldb + n];
}

// C[M,N] (+)= op(A)[M,K] * op(B)[K,N] over k in [k_begin, k_end) for blockIdx.z's split.
template <int BN, bool TA, bool TB>
__global__ void __launch_bounds__(GEMM_THREADS)
    gemm_tile_kernel(int64_t M, int32_t N, int64_t K, const float *__restrict__ A, int64_t lda,
                     const float *__restrict__ B, int64_t ldb, float *__restrict__ C, int64_t ldc, int64_t k_chunk,
                     int64_t split_stride, GemmEpilogue ep) {
    constexpr int TM = 8;
    constexpr int TN = BN / 16;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int32_t n0 = blockIdx.y * BN;
    const int64_t k_begin = (int64_t)blockIdx.z * k_chunk;
    const int64_t k_end = min(K, k_begin + k_chunk);
    float *Cz = C + (int64_t)blockIdx.z * split_stride;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

    // cooperative tile loads: each thread fetches BM*BK/256 = 8 elements of A and BN*BK/256 of B
    constexpr int A_PER = BM * BK / GEMM_THREADS;
    constexpr int B_PER = BN * BK / GEMM_THREADS;
    float ra[A_PER], rb[B_PER];

    auto fetch = [&](int64_t k0) {
#pragma unroll
        for (int i = 0; i < A_PER; i++) {
            const int e = tid + i * GEMM_THREADS;
            int mm, kk;
            if (TA) { mm = e % BM; kk = e / BM; }      // m contiguous in memory
            else { kk = e % BK; mm = e / BK; }         // k contiguous in memory
            const int64_t gm = m0 + mm, gk = k0 + kk;
            ra[i] = (gm < M && gk < k_end) ? ldA<TA>(A, lda, gm, gk) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < B_PER; i++) {
            const int e = tid + i * GEMM_THREADS;
            int nn, kk;
            if (TB) { kk = e % BK; nn = e / BK; }      // k contiguous in memory
            else { nn = e % BN; kk = e / BN; }         // n contiguous in memory
            const int64_t gn = n0 + nn, gk = k0 + kk;
            rb[i] = (gn < N && gk < k_end) ? ldB<TB>(B, ldb, gk, gn) : 0.f;
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; i++) {
            const int e = tid + i * GEMM_THREADS;
            int mm, kk;
            if (TA) { mm = e % BM; kk = e / BM; }
            else { kk = e % BK; mm = e / BK; }
            As[buf][kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_PER; i++) {
            const int e = tid + i * GEMM_THREADS;
            int nn, kk;
            if (TB) { kk = e % BK; nn = e / BK; }
            else { nn = e % BN; kk = e / BN; }
            Bs[buf][kk][nn] = rb[i];
        }
    };

    int buf = 0;
    if (k_begin < k_end) {
        fetch(k_begin);
        stage(0);
    }
    __syncthreads();
    for (int64_t k0 = k_begin; k0 < k_end; k0 += BK) {
        const bool has_next = k0 + BK < k_end;
        if (has_next) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float a[TM], b[TN];
            // rows: ty*4..+4 and 64+ty*4..+4 ; cols: (TN==8) tx*4..+4 and BN/2+tx*4..+4, else tx*TN..+TN
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][kk][64 + ty * 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            if constexpr (TN == 8) {
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][BN / 2 + tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
            } else if constexpr (TN == 4) {
                const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][kk][tx * 4]);
                b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            } else {
                const float2 b0 = *reinterpret_cast<const float2 *>(&Bs[buf][kk][tx * 2]);
                b[0] = b0.x; b[1] = b0.y;
            }
#pragma unroll
            for (int i = 0; i < TM; i++)
#pragma unroll
                for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) {
            stage(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }

#pragma unroll
    for (int i = 0; i < TM; i++) {
        const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; j++) {
            int32_t gn;
            if constexpr (TN == 8) gn = n0 + (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4));
            else gn = n0 + tx * TN + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (ep.bias) v += ep.bias[gn];
            if (ep.relu) v = v > 0.f ? v : 0.f;
            if (ep.mask) v = ep.mask[gm * ep.ldm + gn] > 0.f ? v : 0.f;
            Cz[gm * ldc + gn] = v;
        }
    }
}

// C[i] = sum_{z ascending} partial[z][i]   (fixed order -> deterministic)
__global__ void splitk_reduce_kernel(const float *__restrict__ partial, int64_t n, int splits, int64_t stride,
                                     int32_t ncols, float *__restrict__ C, int64_t ldc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < splits; z++) s += partial[(int64_t)z * stride + i];
    C[(i / ncols) * ldc + (i % ncols)] = s;
}

template <bool TA, bool TB>
static int gemm_dispatch(gnn_ctx *ctx, int64_t M, int32_t N, int64_t K, const float *A, int64_t lda, const float *B,
                         int64_t ldb, float *C, int64_t ldc, GemmEpilogue ep, int splits) {
    const int BN = N > 64 ? 128 : (N > 32 ? 64 : 32);
    dim3 grid((unsigned)ceil_div(M, BM), (unsigned)ceil_div(N, BN), (unsigned)splits);
    float *out = C;
    int64_t out_ld = ldc, stride = 0, k_chunk = K;
    if (splits > 1) {
        k_chunk = round_up(ceil_div(K, splits), BK);
        stride = M * N;
        void *ws = nullptr;
        GNN_TRY(ctx->workspace((size_t)splits * stride * 4, &ws));
        out = (float *)ws;
        out_ld = N;
    }
#define LAUNCH(BNV)                                                                                          \
    gemm_tile_kernel<BNV, TA, TB><<<grid, GEMM_THREADS, 0, ctx->stream>>>(M, N, K, A, lda, B, ldb, out, out_ld, \
                                                                         k_chunk, stride, ep)
    if (BN == 128) LAUNCH(128);
    else if (BN == 64) LAUNCH(64);
    else LAUNCH(32);
#undef LAUNCH
    GNN_LAUNCHED(ctx);
    if (splits > 1) {
        splitk_reduce_kernel<<<(unsigned)ceil_div(stride, 256), 256, 0, ctx->stream>>>(out, stride, splits, stride, N,
                                                                                     C, ldc);
        GNN_LAUNCHED(ctx);
    }
    return 0;
}

// tensor-core path (gemm_tc.cu); returns -1 when the shape is not supported so the caller falls back
int gemm_tc_nt(gnn_ctx *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B, int64_t ldb,
               float *C, int64_t ldc, const float *bias, int relu);
int gemm_tc_nn(gnn_ctx *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B, int64_t ldb,
               float *C, int64_t ldc, const float *mask, int64_t ldm, float *colsum_out = nullptr);
int gemm_tc_tn(gnn_ctx *ctx, int64_t M, int32_t K1, int32_t K2, const float *A, int64_t lda, const float *B,
               int64_t ldb, float *C, int64_t ldc);

// dH = (dP W) . [mask > 0] as gnn_gemm_nn, plus db = column sums of the result (the bias gradient of the layer that
// consumes dH as its dZ): fused into the tensor-core epilogue when the shape allows (*fused = true), otherwise the plain
// product and the caller sums the columns itself.
int gemm_nn_bias_grad(gnn_ctx *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B,
                      int64_t ldb, float *C, int64_t ldc, const float *mask, int64_t ldm, int precision, float *db,
                      bool *fused) {
    *fused = false;
    const char *e = getenv("GNN_FUSED_BIAS_GRAD"); // =0: separate column-sum kernel (A/B runs, tests)
    const bool enabled = !e || atoi(e) != 0;
    if (enabled && precision == 1 && db && M > 0 && N > 0 && K > 0) {
        const int r = gemm_tc_nn(ctx, M, N, K, A, lda, B, ldb, C, ldc, mask, ldm, db);
        if (r == 0) {
            *fused = true;
            return 0;
        }
        if (r > 0) return r;
    }
    return gnn_gemm_nn(ctx, M, N, K, A, lda, B, ldb, C, ldc, mask, ldm, precision);
}

} // namespace gnn

using namespace gnn;

extern "C" {

int gnn_gemm_nt(gnn_ctx_t *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B,
                int64_t ldb, float *C, int64_t ldc, const float *bias, int relu, int precision) {
    GNN_REQUIRE(ctx && A && B && C, "gnn_gemm_nt: NULL argument");
    GNN_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldb >= K && ldc >= N,
                "tensors are not compatible, tensors should of shape [...,A,B] and [...,B,A]");
    if (precision == 1) {
        int r = gemm_tc_nt(ctx, M, N, K, A, lda, B, ldb, C, ldc, bias, relu);
        if (r >= 0) return r;
    }
    GemmEpilogue ep{bias, relu, nullptr, 0};
    return gemm_dispatch<false, true>(ctx, M, N, K, A, lda, B, ldb, C, ldc, ep, 1);
}

int gnn_gemm_nn(gnn_ctx_t *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B,
                int64_t ldb, float *C, int64_t ldc, const float *mask, int64_t ldm, int precision) {
    GNN_REQUIRE(ctx && A && B && C, "gnn_gemm_nn: NULL argument");
    GNN_REQUIRE(M > 0 && N > 0 && K > 0 && lda >= K && ldb >= N && ldc >= N,
                "tensors are not compatible, tensors should of shape [...,A,B] and [...,B,A]");
    if (precision == 1) {
        int r = gemm_tc_nn(ctx, M, N, K, A, lda, B, ldb, C, ldc, mask, ldm);
        if (r >= 0) return r;
    }
    GemmEpilogue ep{nullptr, 0, mask, ldm};
    return gemm_dispatch<false, false>(ctx, M, N, K, A, lda, B, ldb, C, ldc, ep, 1);
}

int gnn_gemm_tn(gnn_ctx_t *ctx, int64_t M, int32_t K1, int32_t K2, const float *A, int64_t lda, const float *B,
                int64_t ldb, float *C, int64_t ldc, int precision) {
    GNN_REQUIRE(ctx && A && B && C, "gnn_gemm_tn: NULL argument");
    GNN_REQUIRE(M > 0 && K1 > 0 && K2 > 0 && lda >= K1 && ldb >= K2 && ldc >= K2,
                "tensors are not compatible, tensors should of shape [...,A,B] and [...,B,A]");
    if (precision == 1) {
        int r = gemm_tc_tn(ctx, M, K1, K2, A, lda, B, ldb, C, ldc);
        if (r >= 0) return r;
    }
    // output is [K1,K2] (small), reduction over the M node rows (huge): split it over ~2 waves of CTAs
    const int BN = K2 > 64 ? 128 : (K2 > 32 ? 64 : 32);
    const int64_t tiles = ceil_div(K1, BM) * ceil_div(K2, BN);
    int64_t splits = ceil_div((int64_t)ctx->sm_count * 2, tiles);
    const int64_t max_splits = ceil_div(M, 4 * BK);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    GemmEpilogue ep{nullptr, 0, nullptr, 0};
    // op(A)[K1, M] = A^T, op(B)[M, K2] = B: "M" of the kernel is K1, "K" is the node dimension
    return gemm_dispatch<true, false>(ctx, K1, K2, M, A, lda, B, ldb, C, ldc, ep, (int)splits);
}

} // extern "C"
