// elementwise.cu — epilogue, loss, optimiser and cyg::tensor-surface kernels (K7, K8, K9).
// All HBM-bound streaming kernels: grid-stride over whole waves of the SMs, fixed-order reductions
// (per-block partials combined by one block) so results are deterministic run to run.
#include <math.h>

#include "common.cuh"

namespace gnn {

static inline unsigned stream_grid(gnn_ctx *ctx, int64_t n, int threads, int per_thread = 4) {
    int64_t want = ceil_div(n, (int64_t)threads * per_thread);
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (want < 1) want = 1;
    return (unsigned)(want < cap ? want : cap);
}

__global__ void fill_kernel(float *__restrict__ p, float v, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void bias_relu_kernel(int64_t N, int32_t F, const float *__restrict__ Y, int64_t ldy,
                                 const float *__restrict__ bias, int relu, float *__restrict__ out, int64_t ldo) {
    const int64_t total = N * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / F;
        const int32_t c = (int32_t)(i - r * F);
        float v = Y[r * ldy + c];
        if (bias) v += bias[c];
        if (relu) v = v > 0.f ? v : 0.f;
        out[r * ldo + c] = v;
    }
}

__global__ void relu_bwd_kernel(int64_t N, int32_t F, const float *__restrict__ dH, int64_t ldd,
                                const float *__restrict__ act, int64_t lda, float *__restrict__ dZ, int64_t ldo) {
    const int64_t total = N * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / F;
        const int32_t c = (int32_t)(i - r * F);
        dZ[r * ldo + c] = act[r * lda + c] > 0.f ? dH[r * ldd + c] : 0.f;
    }
}

// column sums, stage 1: block b sums rows [b*rows_per_block, ...) for every column (threads over columns,
// coalesced across a row); stage 2: one block adds the per-block partials in ascending block order.
constexpr int CS_THREADS = 256;
__global__ void __launch_bounds__(CS_THREADS) colsum_partial_kernel(int64_t N, int32_t F, const float *__restrict__ A,
                                                                    int64_t lda, int64_t rows_per_block,
                                                                    float *__restrict__ partial) {
    // 2-D thread layout: tx over columns (32 wide), ty over rows (8 deep); smem combine over ty in fixed order
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(N, r0 + rows_per_block);
    for (int32_t c0 = 0; c0 < F; c0 += 32) {
        const int32_t c = c0 + tx;
        float s = 0.f;
        if (c < F)
            for (int64_t r = r0 + ty; r < r1; r += 8) s += A[r * lda + c];
        red[ty][tx] = s;
        __syncthreads();
        if (ty == 0 && c < F) {
            float t = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) t += red[j][tx];
            partial[(int64_t)blockIdx.x * F + c] = t;
        }
        __syncthreads();
    }
}
// vector variant (F % 4 == 0, 16-byte aligned rows): a thread owns one float4 column group, 4 independent row loads
// in flight; rows are combined over the thread's row lanes in fixed order through shared memory
template <int TX> // threads across a row (F/4 <= TX), CS_THREADS/TX row lanes
__global__ void __launch_bounds__(CS_THREADS)
    colsum_partial_vec_kernel(int64_t N, int32_t nvec, const float *__restrict__ A, int64_t lda, int64_t rows_per_block,
                              float *__restrict__ partial) {
    constexpr int TY = CS_THREADS / TX;
    __shared__ float4 red[TY][TX];
    const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(N, r0 + rows_per_block);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tx < nvec) {
        int64_t r = r0 + ty;
        for (; r + 3 * TY < r1; r += 4 * TY) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(A + r * lda) + tx);
            const float4 b = __ldg(reinterpret_cast<const float4 *>(A + (r + TY) * lda) + tx);
            const float4 c = __ldg(reinterpret_cast<const float4 *>(A + (r + 2 * TY) * lda) + tx);
            const float4 d = __ldg(reinterpret_cast<const float4 *>(A + (r + 3 * TY) * lda) + tx);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
            s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w;
            s.x += c.x; s.y += c.y; s.z += c.z; s.w += c.w;
            s.x += d.x; s.y += d.y; s.z += d.z; s.w += d.w;
        }
        for (; r < r1; r += TY) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(A + r * lda) + tx);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && tx < nvec) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < TY; j++) { t.x += red[j][tx].x; t.y += red[j][tx].y; t.z += red[j][tx].z; t.w += red[j][tx].w; }
        reinterpret_cast<float4 *>(partial + (int64_t)blockIdx.x * nvec * 4)[tx] = t;
    }
}
__global__ void colsum_final_kernel(int32_t F, int nblocks, const float *__restrict__ partial, float *__restrict__ out,
                                    int32_t ldp) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= F) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; b++) s += partial[(int64_t)b * ldp + c];
    out[c] = s;
}

// softmax cross-entropy: one warp per row, classes strided over lanes.
constexpr int XE_THREADS = 256;
__global__ void __launch_bounds__(XE_THREADS)
    softmax_xent_kernel(int64_t N, int32_t C, const float *__restrict__ Z, int64_t ldz, const int32_t *__restrict__ y,
                        float inv_n, float *__restrict__ dZ, int64_t ldd, float *__restrict__ partial) {
    __shared__ float wsum[XE_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * (XE_THREADS / 32);
    float local = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * (XE_THREADS / 32) + warp; r < N; r += nwarps) {
        const float *z = Z + r * ldz;
        float m = -INFINITY;
        for (int32_t c = lane; c < C; c += 32) m = fmaxf(m, z[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int32_t c = lane; c < C; c += 32) s += expf(z[c] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const int32_t yi = y[r];
        const float zy = z[yi];
        // reference: -log( exp(z_y) / (sum exp(z) + 1e-20) )  (src/nn.cpp:446-450), evaluated with the row
        // max factored out so it cannot overflow: exp(z_y-m) / (sum exp(z-m) + 1e-20*exp(-m))
        const float eps = (m > -60.f) ? 1e-20f * expf(-m) : INFINITY;
        const float li = -logf(expf(zy - m) / (s + eps));
        if (lane == 0) local += li;
        if (dZ) {
            const float inv_s = 1.f / s;
            for (int32_t c = lane; c < C; c += 32) {
                float p = expf(z[c] - m) * inv_s;       // nn::softmax (src/nn.cpp:270-278)
                if (c == yi) p -= 1.f;
                dZ[r * ldd + c] = p * inv_n;
            }
        }
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < XE_THREADS / 32; w++) t += wsum[w];
        partial[blockIdx.x] = t;
    }
}
// Tiled variant for C <= 64: persistent CTAs stage 128 rows at a time through shared memory with coalesced 128-bit
// loads, one thread then owns one row (max, sum of exp, loss term, dZ written back into the tile), the tile leaves
// with coalesced stores, and thread c < C adds column c of the tile to its running bias-gradient sum.  Per-CTA
// loss and column-sum partials are combined by fixed-order final kernels.
constexpr int XT_ROWS = 128, XT_THREADS = 128;
__global__ void __launch_bounds__(XT_THREADS)
    softmax_xent_tile_kernel(int64_t N, int32_t C, const float *__restrict__ Z, int64_t ldz, const int32_t *__restrict__ y,
                             float inv_n, float *__restrict__ dZ, int64_t ldd, float *__restrict__ loss_partial,
                             float *__restrict__ db_partial, int32_t ldw /* padded C, multiple of 4 */) {
    extern __shared__ float tile[]; // [XT_ROWS][ldw + 1]
    const int32_t lds = ldw + 1;
    const int32_t nvec = ldw / 4;
    float loss_acc = 0.f, db_acc = 0.f;
    const int64_t n_tiles = (N + XT_ROWS - 1) / XT_ROWS;
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t r0 = t * XT_ROWS;
        const int32_t rows = (int32_t)min((int64_t)XT_ROWS, N - r0);
        for (int32_t i = threadIdx.x; i < rows * nvec; i += XT_THREADS) {
            const int32_t r = i / nvec, v = i - r * nvec;
            const float4 a = __ldg(reinterpret_cast<const float4 *>(Z + (r0 + r) * ldz) + v);
            float *d = tile + r * lds + v * 4;
            d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w;
        }
        __syncthreads();
        if ((int32_t)threadIdx.x < rows) {
            float *z = tile + threadIdx.x * lds;
            float mx = -INFINITY;
            for (int32_t c = 0; c < C; c++) mx = fmaxf(mx, z[c]);
            float s = 0.f;
            for (int32_t c = 0; c < C; c++) s += expf(z[c] - mx);
            const int32_t yi = y[r0 + threadIdx.x];
            const float zy = z[yi];
            // same evaluation as softmax_xent_kernel (reference formula, src/nn.cpp:446-450, max factored out)
            const float eps = (mx > -60.f) ? 1e-20f * expf(-mx) : INFINITY;
            loss_acc += -logf(expf(zy - mx) / (s + eps));
            const float inv_s = 1.f / s;
            for (int32_t c = 0; c < C; c++) {
                float pr = expf(z[c] - mx) * inv_s;
                if (c == yi) pr -= 1.f;
                z[c] = pr * inv_n;
            }
            for (int32_t c = C; c < ldw; c++) z[c] = 0.f;
        }
        __syncthreads();
        if (dZ)
            for (int32_t i = threadIdx.x; i < rows * nvec; i += XT_THREADS) {
                const int32_t r = i / nvec, v = i - r * nvec;
                const float *d = tile + r * lds + v * 4;
                reinterpret_cast<float4 *>(dZ + (r0 + r) * ldd)[v] = make_float4(d[0], d[1], d[2], d[3]);
            }
        if (db_partial && (int32_t)threadIdx.x < C) {
            float cs = 0.f;
            for (int32_t r = 0; r < rows; r++) cs += tile[r * lds + threadIdx.x];
            db_acc += cs;
        }
        __syncthreads();
    }
    // loss: fixed-order sum over the CTA's threads
    __shared__ float red[XT_THREADS];
    red[threadIdx.x] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
        for (int i = 0; i < XT_THREADS; i++) tsum += red[i];
        loss_partial[blockIdx.x] = tsum;
    }
    if (db_partial && (int32_t)threadIdx.x < C) db_partial[(int64_t)blockIdx.x * C + threadIdx.x] = db_acc;
}
__global__ void xent_final_kernel(int nblocks, const float *__restrict__ partial, float inv_n, float *__restrict__ loss) {
    // one warp, fixed order: lane-strided partial sums then a shuffle tree
    float s = 0.f;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *loss = s * inv_n;
}

__global__ void sgd_kernel(int64_t n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ vel,
                           float lr, float momentum, float dampening, float wd, int nesterov, int first) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float d = g[i];
        const float pi = p[i];
        if (wd != 0.f) d = d + wd * pi;
        if (momentum != 0.f) {
            const float v = first ? d : momentum * vel[i] + (1.f - dampening) * d;
            vel[i] = v;
            d = nesterov ? d + momentum * v : v;
        }
        p[i] = pi - lr * d;
    }
}

// torch.optim.Adam semantics (the intent of nn::Adam, reference include/nn.h:180-188; the body, src/nn.cpp:419-441,
// divides by sqrt(v)*eps and uses the parameter index as the step count).  bc1 = 1 - b1^t, bc2 = 1 - b2^t.
__global__ void adam_kernel(int64_t n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                            float *__restrict__ v, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        float d = g[i];
        if (wd != 0.f) d = d + wd * pi;
        const float mi = b1 * m[i] + (1.f - b1) * d;
        const float vi = b2 * v[i] + (1.f - b2) * d * d;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

// masked softmax cross-entropy: one warp per row like softmax_xent_kernel, rows with mask[r] == 0 contribute neither
// to the loss nor to dZ (their dZ row is zero); the mean is over n_sel (selected rows of the whole graph)
__global__ void __launch_bounds__(256)
    softmax_xent_masked_kernel(int64_t N, int32_t C, const float *__restrict__ Z, int64_t ldz, const int32_t *__restrict__ y,
                               const uint8_t *__restrict__ mask, float inv_n, float *__restrict__ dZ, int64_t ldd,
                               float *__restrict__ partial) {
    __shared__ float wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    float local = 0.f;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < N; r += nwarps) {
        if (!mask[r]) {
            if (dZ)
                for (int32_t c = lane; c < C; c += 32) dZ[r * ldd + c] = 0.f;
            continue;
        }
        const float *z = Z + r * ldz;
        float mx = -INFINITY;
        for (int32_t c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float s = 0.f;
        for (int32_t c = lane; c < C; c += 32) s += expf(z[c] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const int32_t yi = y[r];
        const float eps = (mx > -60.f) ? 1e-20f * expf(-mx) : INFINITY;
        const float li = -logf(expf(z[yi] - mx) / (s + eps));
        if (lane == 0) local += li;
        if (dZ) {
            const float inv_s = 1.f / s;
            for (int32_t c = lane; c < C; c += 32) {
                float pr = expf(z[c] - mx) * inv_s;
                if (c == yi) pr -= 1.f;
                dZ[r * ldd + c] = pr * inv_n;
            }
        }
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; w++) t += wsum[w];
        partial[blockIdx.x] = t;
    }
}

// rows whose arg-max logit (first maximum, like tensor::argmax, reference include/tensor.h:645-648) equals the label,
// counted over the rows selected by mask (all rows when mask == NULL): exact integer partials, fixed-order sum
__global__ void __launch_bounds__(256)
    argmax_correct_kernel(int64_t N, int32_t C, const float *__restrict__ Z, int64_t ldz, const int32_t *__restrict__ y,
                          const uint8_t *__restrict__ mask, int32_t *__restrict__ partial) {
    __shared__ int32_t wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    int32_t local = 0;
    for (int64_t r = (int64_t)blockIdx.x * 8 + warp; r < N; r += nwarps) {
        if (mask && !mask[r]) continue;
        const float *z = Z + r * ldz;
        float best = -INFINITY;
        int32_t bi = 0x7fffffff;
        for (int32_t c = lane; c < C; c += 32) {
            const float v = z[c];
            if (v > best) { best = v; bi = c; } // strict: keeps the first maximum of this lane's columns
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0 && bi == y[r]) local++;
    }
    if (lane == 0) wsum[warp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t t = 0;
        for (int w = 0; w < 8; w++) t += wsum[w];
        partial[blockIdx.x] = t;
    }
}
__global__ void count_final_kernel(int nblocks, const int32_t *__restrict__ partial, int64_t *__restrict__ out) {
    int64_t s = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partial[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *out = s;
}

__global__ void binary_kernel(int op, int64_t rows, int64_t cols, const float *__restrict__ a, int64_t a_rs,
                              int64_t a_cs, const float *__restrict__ b, int64_t b_rs, int64_t b_cs,
                              float *__restrict__ out) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i - r * cols;
        const float x = a[r * a_rs + c * a_cs], y = b[r * b_rs + c * b_cs];
        float v;
        switch (op) {
        case GNN_OP_ADD: v = x + y; break;
        case GNN_OP_MUL: v = x * y; break;
        case GNN_OP_DIV: v = x / y; break;
        case GNN_OP_POW: v = powf(x, y); break;
        default: v = x > y ? 1.f : 0.f; break;
        }
        out[i] = v;
    }
}
__global__ void unary_kernel(int op, int64_t n, const float *__restrict__ a, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float x = a[i];
        out[i] = op == GNN_UOP_EXP ? expf(x) : (op == GNN_UOP_LOG ? logf(x) : -x);
    }
}
__global__ void where_kernel(int64_t n, const float *__restrict__ cond, const float *__restrict__ t,
                             const float *__restrict__ f, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = cond[i] > 0.f ? t[i] : f[i];
}
// row sums: one warp per row
__global__ void rowsum_kernel(int64_t rows, int64_t cols, const float *__restrict__ a, float *__restrict__ out) {
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= rows) return;
    float s = 0.f;
    for (int64_t c = lane; c < cols; c += 32) s += a[w * cols + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[w] = s;
}
__global__ void transpose_kernel(int64_t rows, int64_t cols, int64_t tiles_c, const float *__restrict__ a,
                                 float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int64_t c0 = ((int64_t)blockIdx.x % tiles_c) * 32, r0 = ((int64_t)blockIdx.x / tiles_c) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < rows && c < cols) tile[j][threadIdx.x] = a[r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t c = c0 + j, r = r0 + threadIdx.x;
        if (r < rows && c < cols) out[c * rows + r] = tile[threadIdx.x][j];
    }
}
__global__ void gather_cols_kernel(int64_t rows, int64_t cols, const float *__restrict__ a,
                                   const int32_t *__restrict__ idx, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) out[i] = a[i * cols + idx[i]];
}

// dst[r, 0:cols] = src[r, 0:cols] for two row-major views (a strided cudaMemcpy2DAsync of sub-kilobyte rows runs far
// below HBM speed; this moves 128-bit vectors when the views allow it)
__global__ void copy2d_kernel(float *__restrict__ dst, int64_t ldd, const float *__restrict__ src, int64_t lds,
                              int64_t rows, int32_t cols, int vec) {
    const int32_t cpr = vec ? cols / 4 : cols; // elements per row in units of the access width
    const int64_t total = rows * cpr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cpr;
        const int32_t c = (int32_t)(i - r * cpr);
        if (vec) reinterpret_cast<float4 *>(dst + r * ldd)[c] = __ldg(reinterpret_cast<const float4 *>(src + r * lds) + c);
        else dst[r * ldd + c] = src[r * lds + c];
    }
}

int copy2d(gnn_ctx *ctx, float *dst, int64_t ldd, const float *src, int64_t lds, int64_t rows, int32_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    if (ldd == cols && lds == cols) {
        GNN_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)rows * cols * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    const int vec = (cols % 4 == 0) && (ldd % 4 == 0) && (lds % 4 == 0) && ((uintptr_t)dst % 16 == 0) && ((uintptr_t)src % 16 == 0);
    copy2d_kernel<<<stream_grid(ctx, rows * (vec ? cols / 4 : cols), 256), 256, 0, ctx->stream>>>(dst, ldd, src, lds, rows, cols, vec);
    GNN_LAUNCHED(ctx);
    return 0;
}

// ---- BatchNorm over the node dimension (nn::BatchNorm, reference src/nn.cpp:285-330) ----------------------------
// column statistics: block b reduces rows [b*rows_per_block, ...) of up to three per-column quantities
//   MODE 0: sum x                                   (mean)
//   MODE 1: sum (x - mean)^2                        (two-pass variance like functional::var, functional.h:383-387)
//   MODE 2: sum g and sum g * xhat, g = dY masked by the forward output (backward reductions)
template <int MODE>
__global__ void __launch_bounds__(CS_THREADS)
    bn_partial_kernel(int64_t N, int32_t F, const float *__restrict__ X, int64_t ldx, const float *__restrict__ mean,
                      const float *__restrict__ var, float eps, const float *__restrict__ dY, int64_t ldd,
                      const float *__restrict__ Yout, int64_t ldy, int64_t rows_per_block, float *__restrict__ partial) {
    __shared__ float red[2][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(N, r0 + rows_per_block);
    for (int32_t c0 = 0; c0 < F; c0 += 32) {
        const int32_t c = c0 + tx;
        float s0 = 0.f, s1 = 0.f;
        if (c < F) {
            const float m = MODE >= 1 ? mean[c] : 0.f;
            const float istd = MODE == 2 ? rsqrtf(var[c] + eps) : 0.f;
            for (int64_t r = r0 + ty; r < r1; r += 8) {
                const float x = X[r * ldx + c];
                if (MODE == 0) s0 += x;
                else if (MODE == 1) s0 += (x - m) * (x - m);
                else {
                    float g = dY[r * ldd + c];
                    if (Yout && !(Yout[r * ldy + c] > 0.f)) g = 0.f;
                    s0 += g;
                    s1 += g * ((x - m) * istd);
                }
            }
        }
        red[0][ty][tx] = s0;
        red[1][ty][tx] = s1;
        __syncthreads();
        if (ty == 0 && c < F) {
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) { t0 += red[0][j][tx]; t1 += red[1][j][tx]; }
            partial[(int64_t)blockIdx.x * 2 * F + c] = t0;
            partial[(int64_t)blockIdx.x * 2 * F + F + c] = t1;
        }
        __syncthreads();
    }
}
// out[c] = scale * sum_b partial[b][which][c]   (fixed block order)
__global__ void bn_final_kernel(int32_t F, int nblocks, const float *__restrict__ partial, int which, float scale,
                                float *__restrict__ out) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= F) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; b++) s += partial[(int64_t)b * 2 * F + which * F + c];
    out[c] = s * scale;
}
__global__ void bn_apply_kernel(int64_t N, int32_t F, const float *__restrict__ X, int64_t ldx, const float *__restrict__ mean,
                                const float *__restrict__ var, float eps, const float *__restrict__ gamma,
                                const float *__restrict__ beta, int relu, float *__restrict__ Y, int64_t ldy) {
    const int64_t total = N * F;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / F;
        const int32_t c = (int32_t)(i - r * F);
        float v = (X[r * ldx + c] - mean[c]) / sqrtf(var[c] + eps);
        v = v * gamma[c];
        if (beta) v += beta[c];
        if (relu) v = v > 0.f ? v : 0.f;
        Y[r * ldy + c] = v;
    }
}
// dX = gamma * istd * (g - mean(g) - xhat * mean(g xhat))
__global__ void bn_bwd_apply_kernel(int64_t N, int32_t F, const float *__restrict__ X, int64_t ldx,
                                    const float *__restrict__ mean, const float *__restrict__ var, float eps,
                                    const float *__restrict__ gamma, const float *__restrict__ dY, int64_t ldd,
                                    const float *__restrict__ Yout, int64_t ldy, const float *__restrict__ sum_g,
                                    const float *__restrict__ sum_gx, float *__restrict__ dX, int64_t ldo) {
    const int64_t total = N * F;
    const float inv_n = 1.f / (float)N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / F;
        const int32_t c = (int32_t)(i - r * F);
        const float istd = rsqrtf(var[c] + eps);
        float g = dY[r * ldd + c];
        if (Yout && !(Yout[r * ldy + c] > 0.f)) g = 0.f;
        const float xh = (X[r * ldx + c] - mean[c]) * istd;
        dX[r * ldo + c] = gamma[c] * istd * (g - sum_g[c] * inv_n - xh * sum_gx[c] * inv_n);
    }
}

// ---- LayerNorm over the feature dimension (nn::LayerNorm, reference src/nn.cpp:332-353): one warp per row --------
__global__ void __launch_bounds__(256)
    layernorm_fwd_kernel(int64_t N, int32_t F, const float *__restrict__ X, int64_t ldx, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float eps, int relu, float *__restrict__ Y, int64_t ldy,
                         float *__restrict__ mean, float *__restrict__ rstd) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const float *x = X + r * ldx;
    float s = 0.f;
    for (int32_t c = lane; c < F; c += 32) s += x[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float m = s / (float)F;
    float q = 0.f;
    for (int32_t c = lane; c < F; c += 32) { const float d = x[c] - m; q += d * d; } // two passes like functional::var
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = 1.0f / sqrtf(q / (float)F + eps);
    if (lane == 0) { mean[r] = m; rstd[r] = rs; }
    for (int32_t c = lane; c < F; c += 32) {
        float v = (x[c] - m) * rs;
        if (gamma) v *= gamma[c];
        if (beta) v += beta[c];
        if (relu) v = v > 0.f ? v : 0.f;
        Y[r * ldy + c] = v;
    }
}
// dX = rstd * (gg - mean(gg) - xhat * mean(gg xhat)),  gg = g * gamma, g = dY masked by the forward output
__global__ void __launch_bounds__(256)
    layernorm_bwd_kernel(int64_t N, int32_t F, const float *__restrict__ X, int64_t ldx, const float *__restrict__ mean,
                         const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ Yout,
                         int64_t ldy, const float *__restrict__ dY, int64_t ldd, float *__restrict__ dX, int64_t ldo) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const float m = mean[r], rs = rstd[r];
    float a = 0.f, b = 0.f;
    for (int32_t c = lane; c < F; c += 32) {
        float g = dY[r * ldd + c];
        if (Yout && !(Yout[r * ldy + c] > 0.f)) g = 0.f;
        const float gg = g * (gamma ? gamma[c] : 1.f);
        const float xh = (X[r * ldx + c] - m) * rs;
        a += gg;
        b += gg * xh;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    const float inv_f = 1.f / (float)F;
    for (int32_t c = lane; c < F; c += 32) {
        float g = dY[r * ldd + c];
        if (Yout && !(Yout[r * ldy + c] > 0.f)) g = 0.f;
        const float gg = g * (gamma ? gamma[c] : 1.f);
        const float xh = (X[r * ldx + c] - m) * rs;
        dX[r * ldo + c] = rs * (gg - a * inv_f - xh * b * inv_f);
    }
}
// per-column sums of g and g * xhat with PER-ROW statistics (dbeta, dgamma of LayerNorm): same layout as bn_partial
__global__ void __launch_bounds__(CS_THREADS)
    ln_colred_partial_kernel(int64_t N, int32_t F, const float *__restrict__ X, int64_t ldx, const float *__restrict__ mean,
                             const float *__restrict__ rstd, const float *__restrict__ dY, int64_t ldd,
                             const float *__restrict__ Yout, int64_t ldy, int64_t rows_per_block, float *__restrict__ partial) {
    __shared__ float red[2][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t r1 = min(N, r0 + rows_per_block);
    for (int32_t c0 = 0; c0 < F; c0 += 32) {
        const int32_t c = c0 + tx;
        float s0 = 0.f, s1 = 0.f;
        if (c < F)
            for (int64_t r = r0 + ty; r < r1; r += 8) {
                float g = dY[r * ldd + c];
                if (Yout && !(Yout[r * ldy + c] > 0.f)) g = 0.f;
                s0 += g;
                s1 += g * ((X[r * ldx + c] - mean[r]) * rstd[r]);
            }
        red[0][ty][tx] = s0;
        red[1][ty][tx] = s1;
        __syncthreads();
        if (ty == 0 && c < F) {
            float t0 = 0.f, t1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) { t0 += red[0][j][tx]; t1 += red[1][j][tx]; }
            partial[(int64_t)blockIdx.x * 2 * F + c] = t0;
            partial[(int64_t)blockIdx.x * 2 * F + F + c] = t1;
        }
        __syncthreads();
    }
}
__global__ void tanh_fwd_kernel(int64_t n, const float *__restrict__ x, float *__restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = tanhf(x[i] + 1e-12f); // = (e^s - e^-s)/(e^s + e^-s) of nn::tanh (src/nn.cpp:355-364) without its overflow
}
__global__ void tanh_bwd_kernel(int64_t n, const float *__restrict__ y, const float *__restrict__ dy, float *__restrict__ dx) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * (1.f - y[i] * y[i]);
}
// counter-based keep mask, identical to oracle/gcn_oracle.c:orc_dropout_fwd (splitmix64 of (seed, stream 77, index))
__device__ __forceinline__ uint64_t mix64_dev(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void dropout_kernel(int64_t n, const float *__restrict__ x, float p, uint64_t base, float *__restrict__ y) {
    const float scale = 1.f / (1.f - p);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float u = (float)(mix64_dev(base + (uint64_t)i) >> 40) * (1.0f / 16777216.0f);
        y[i] = u >= p ? x[i] * scale : 0.f;
    }
}

static int bn_blocks(gnn_ctx *ctx, int64_t N, int64_t *rows_per_block) {
    int64_t nblocks = (int64_t)ctx->sm_count * 4;
    int64_t rpb = ceil_div(N, nblocks);
    if (rpb < 64) rpb = 64;
    rpb = round_up(rpb, 8);
    *rows_per_block = rpb;
    return (int)ceil_div(N, rpb);
}

int colsum(gnn_ctx *ctx, int64_t N, int32_t F, const float *A, int64_t lda, float *out) {
    int64_t nblocks = (int64_t)ctx->sm_count * 4;
    int64_t rows_per_block = ceil_div(N, nblocks);
    if (rows_per_block < 64) rows_per_block = 64;
    rows_per_block = round_up(rows_per_block, 8);
    nblocks = ceil_div(N, rows_per_block);
    void *ws = nullptr;
    const int32_t Fp = (int32_t)round_up(F, 4);
    GNN_TRY(ctx->workspace((size_t)nblocks * Fp * 4, &ws));
    // vector path: the padding columns of a row (up to lda) are read and summed into unused partial slots
    const bool vec = ((uintptr_t)A % 16 == 0) && (lda % 4 == 0) && lda >= Fp && Fp <= 512;
    if (vec) {
        const int32_t nvec = Fp / 4;
        if (nvec <= 16)
            colsum_partial_vec_kernel<16><<<(unsigned)nblocks, CS_THREADS, 0, ctx->stream>>>(N, nvec, A, lda, rows_per_block, (float *)ws);
        else if (nvec <= 32)
            colsum_partial_vec_kernel<32><<<(unsigned)nblocks, CS_THREADS, 0, ctx->stream>>>(N, nvec, A, lda, rows_per_block, (float *)ws);
        else if (nvec <= 64)
            colsum_partial_vec_kernel<64><<<(unsigned)nblocks, CS_THREADS, 0, ctx->stream>>>(N, nvec, A, lda, rows_per_block, (float *)ws);
        else
            colsum_partial_vec_kernel<128><<<(unsigned)nblocks, CS_THREADS, 0, ctx->stream>>>(N, nvec, A, lda, rows_per_block, (float *)ws);
        GNN_LAUNCHED(ctx);
        colsum_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, (int)nblocks, (const float *)ws, out, Fp);
        GNN_LAUNCHED(ctx);
        return 0;
    }
    colsum_partial_kernel<<<(unsigned)nblocks, CS_THREADS, 0, ctx->stream>>>(N, F, A, lda, rows_per_block, (float *)ws);
    GNN_LAUNCHED(ctx);
    colsum_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, (int)nblocks, (const float *)ws, out, F);
    GNN_LAUNCHED(ctx);
    return 0;
}

// loss + dZ (+ db = column sums of dZ when db != NULL, which needs the tiled kernel: returns through *db_done)
int softmax_xent_launch(gnn_ctx *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                        int64_t n_total, float *loss, float *dZ, int64_t ldd, float *db, bool may_touch_padding) {
    GNN_REQUIRE(ctx && Z && y && loss, "gnn_softmax_xent: NULL argument");
    GNN_REQUIRE(N > 0 && C > 0 && ldz >= C, "invalid input, logits must be of rank 2 and targets must be 1D tensor");
    if (n_total <= 0) n_total = N;
    const float inv_n = 1.0f / (float)n_total;
    const int32_t ldw = (int32_t)round_up(C, 4);
    const bool tiled = C <= 64 && ((uintptr_t)Z % 16 == 0) && (ldz % 4 == 0) && ldz >= ldw &&
                       (!dZ || (((uintptr_t)dZ % 16 == 0) && (ldd % 4 == 0) && ldd >= ldw &&
                                (may_touch_padding || C % 4 == 0))); // the tile kernel zeroes dZ columns C..ldw-1
    void *ws = nullptr;
    if (tiled) {
        int64_t nblocks = ceil_div(N, XT_ROWS);
        const int64_t cap = (int64_t)ctx->sm_count * 8;
        if (nblocks > cap) nblocks = cap;
        GNN_TRY(ctx->workspace((size_t)nblocks * (C + 1) * 4, &ws));
        float *lp = (float *)ws, *dbp = db ? lp + nblocks : nullptr;
        const size_t smem = (size_t)XT_ROWS * (ldw + 1) * 4;
        softmax_xent_tile_kernel<<<(unsigned)nblocks, XT_THREADS, smem, ctx->stream>>>(N, C, Z, ldz, y, inv_n, dZ, ldd, lp,
                                                                                      dbp, ldw);
        GNN_LAUNCHED(ctx);
        xent_final_kernel<<<1, 32, 0, ctx->stream>>>((int)nblocks, lp, inv_n, loss);
        GNN_LAUNCHED(ctx);
        if (db) {
            colsum_final_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, ctx->stream>>>(C, (int)nblocks, dbp, db, C);
            GNN_LAUNCHED(ctx);
        }
        return 0;
    }
    int64_t nblocks = ceil_div(N, XE_THREADS / 32);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (nblocks > cap) nblocks = cap;
    GNN_TRY(ctx->workspace((size_t)nblocks * 4, &ws));
    softmax_xent_kernel<<<(unsigned)nblocks, XE_THREADS, 0, ctx->stream>>>(N, C, Z, ldz, y, inv_n, dZ, ldd, (float *)ws);
    GNN_LAUNCHED(ctx);
    xent_final_kernel<<<1, 32, 0, ctx->stream>>>((int)nblocks, (const float *)ws, inv_n, loss);
    GNN_LAUNCHED(ctx);
    if (db && dZ) GNN_TRY(colsum(ctx, N, C, dZ, ldd, db));
    return 0;
}

} // namespace gnn

using namespace gnn;

extern "C" {

int gnn_fill_f32(gnn_ctx_t *ctx, float *ptr, float value, int64_t n) {
    GNN_REQUIRE(ctx && (ptr || n == 0), "gnn_fill_f32: NULL argument");
    if (n <= 0) return 0;
    fill_kernel<<<stream_grid(ctx, n, 256), 256, 0, ctx->stream>>>(ptr, value, n);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_bias_relu_fwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *Y, int64_t ldy, const float *bias, int relu,
                      float *out, int64_t ldo) {
    GNN_REQUIRE(ctx && Y && out && N > 0 && F > 0, "gnn_bias_relu_fwd: bad argument");
    bias_relu_kernel<<<stream_grid(ctx, N * F, 256), 256, 0, ctx->stream>>>(N, F, Y, ldy, bias, relu, out, ldo);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_relu_bwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *dH, int64_t ldd, const float *act, int64_t lda,
                 float *dZ, int64_t ldo) {
    GNN_REQUIRE(ctx && dH && act && dZ && N > 0 && F > 0, "gnn_relu_bwd: bad argument");
    relu_bwd_kernel<<<stream_grid(ctx, N * F, 256), 256, 0, ctx->stream>>>(N, F, dH, ldd, act, lda, dZ, ldo);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_bias_grad(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *dZ, int64_t ldd, float *db) {
    GNN_REQUIRE(ctx && dZ && db && N > 0 && F > 0, "gnn_bias_grad: bad argument");
    return colsum(ctx, N, F, dZ, ldd, db);
}

int gnn_softmax_xent(gnn_ctx_t *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                     int64_t n_total, float *loss, float *dZ, int64_t ldd) {
    // a dZ view wider than the padded row (a column slice of a larger matrix) keeps its neighbouring columns
    return softmax_xent_launch(ctx, N, C, Z, ldz, y, n_total, loss, dZ, ldd, nullptr, ldd == round_up(C, 4));
}

int gnn_sgd_step(gnn_ctx_t *ctx, int64_t n, float *p, const float *g, float *vel, float lr, float momentum,
                 float dampening, float weight_decay, int nesterov, int first) {
    GNN_REQUIRE(ctx && p && g && n > 0, "gnn_sgd_step: bad argument");
    GNN_REQUIRE(momentum == 0.f || vel, "gnn_sgd_step: momentum needs a velocity buffer");
    sgd_kernel<<<stream_grid(ctx, n, 256, 1), 256, 0, ctx->stream>>>(n, p, g, vel, lr, momentum, dampening, weight_decay,
                                                                   nesterov, first);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_batchnorm_fwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *gamma,
                      const float *beta, float eps, int relu, float *Y, int64_t ldy, float *mean, float *var) {
    GNN_REQUIRE(ctx && X && gamma && Y && mean && var && N > 0 && F > 0 && ldx >= F && ldy >= F, "gnn_batchnorm_fwd: bad argument");
    int64_t rpb = 0;
    const int nb = bn_blocks(ctx, N, &rpb);
    void *ws = nullptr;
    GNN_TRY(ctx->workspace((size_t)nb * 2 * F * 4, &ws));
    float *part = (float *)ws;
    bn_partial_kernel<0><<<nb, CS_THREADS, 0, ctx->stream>>>(N, F, X, ldx, nullptr, nullptr, eps, nullptr, 0, nullptr, 0, rpb, part);
    GNN_LAUNCHED(ctx);
    bn_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, nb, part, 0, 1.0f / (float)N, mean);
    GNN_LAUNCHED(ctx);
    bn_partial_kernel<1><<<nb, CS_THREADS, 0, ctx->stream>>>(N, F, X, ldx, mean, nullptr, eps, nullptr, 0, nullptr, 0, rpb, part);
    GNN_LAUNCHED(ctx);
    bn_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, nb, part, 0, 1.0f / (float)N, var);
    GNN_LAUNCHED(ctx);
    bn_apply_kernel<<<stream_grid(ctx, N * F, 256), 256, 0, ctx->stream>>>(N, F, X, ldx, mean, var, eps, gamma, beta, relu, Y, ldy);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_batchnorm_bwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *mean, const float *var,
                      const float *gamma, float eps, const float *relu_out, int64_t ldy, const float *dY, int64_t ldd,
                      float *dX, int64_t ldo, float *dgamma, float *dbeta) {
    GNN_REQUIRE(ctx && X && mean && var && gamma && dY && dX && dgamma && dbeta && N > 0 && F > 0, "gnn_batchnorm_bwd: bad argument");
    int64_t rpb = 0;
    const int nb = bn_blocks(ctx, N, &rpb);
    void *ws = nullptr;
    GNN_TRY(ctx->workspace((size_t)nb * 2 * F * 4, &ws));
    float *part = (float *)ws;
    bn_partial_kernel<2><<<nb, CS_THREADS, 0, ctx->stream>>>(N, F, X, ldx, mean, var, eps, dY, ldd, relu_out, ldy, rpb, part);
    GNN_LAUNCHED(ctx);
    bn_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, nb, part, 0, 1.0f, dbeta);
    GNN_LAUNCHED(ctx);
    bn_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, nb, part, 1, 1.0f, dgamma);
    GNN_LAUNCHED(ctx);
    bn_bwd_apply_kernel<<<stream_grid(ctx, N * F, 256), 256, 0, ctx->stream>>>(N, F, X, ldx, mean, var, eps, gamma, dY, ldd,
                                                                              relu_out, ldy, dbeta, dgamma, dX, ldo);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_layernorm_fwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *gamma,
                      const float *beta, float eps, int relu, float *Y, int64_t ldy, float *mean, float *rstd) {
    GNN_REQUIRE(ctx && X && Y && mean && rstd && N > 0 && F > 0 && ldx >= F && ldy >= F, "gnn_layernorm_fwd: bad argument");
    layernorm_fwd_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, ctx->stream>>>(N, F, X, ldx, gamma, beta, eps, relu, Y, ldy,
                                                                                  mean, rstd);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_layernorm_bwd(gnn_ctx_t *ctx, int64_t N, int32_t F, const float *X, int64_t ldx, const float *mean, const float *rstd,
                      const float *gamma, const float *relu_out, int64_t ldy, const float *dY, int64_t ldd, float *dX,
                      int64_t ldo, float *dgamma, float *dbeta) {
    GNN_REQUIRE(ctx && X && mean && rstd && dY && dX && N > 0 && F > 0, "gnn_layernorm_bwd: bad argument");
    layernorm_bwd_kernel<<<(unsigned)ceil_div(N * 32, 256), 256, 0, ctx->stream>>>(N, F, X, ldx, mean, rstd, gamma, relu_out, ldy,
                                                                                  dY, ldd, dX, ldo);
    GNN_LAUNCHED(ctx);
    if (dgamma || dbeta) {
        int64_t rpb = 0;
        const int nb = bn_blocks(ctx, N, &rpb);
        void *ws = nullptr;
        GNN_TRY(ctx->workspace((size_t)nb * 2 * F * 4, &ws));
        float *part = (float *)ws;
        ln_colred_partial_kernel<<<nb, CS_THREADS, 0, ctx->stream>>>(N, F, X, ldx, mean, rstd, dY, ldd, relu_out, ldy, rpb, part);
        GNN_LAUNCHED(ctx);
        if (dbeta) {
            bn_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, nb, part, 0, 1.0f, dbeta);
            GNN_LAUNCHED(ctx);
        }
        if (dgamma) {
            bn_final_kernel<<<(unsigned)ceil_div(F, 128), 128, 0, ctx->stream>>>(F, nb, part, 1, 1.0f, dgamma);
            GNN_LAUNCHED(ctx);
        }
    }
    return 0;
}

int gnn_tanh_fwd(gnn_ctx_t *ctx, int64_t n, const float *x, float *y) {
    GNN_REQUIRE(ctx && x && y && n > 0, "gnn_tanh_fwd: bad argument");
    tanh_fwd_kernel<<<stream_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, x, y);
    GNN_LAUNCHED(ctx);
    return 0;
}
int gnn_tanh_bwd(gnn_ctx_t *ctx, int64_t n, const float *y, const float *dy, float *dx) {
    GNN_REQUIRE(ctx && y && dy && dx && n > 0, "gnn_tanh_bwd: bad argument");
    tanh_bwd_kernel<<<stream_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, y, dy, dx);
    GNN_LAUNCHED(ctx);
    return 0;
}
int gnn_dropout(gnn_ctx_t *ctx, int64_t n, const float *x, float p, uint64_t seed, float *y) {
    GNN_REQUIRE(ctx && x && y && n > 0, "gnn_dropout: bad argument");
    GNN_REQUIRE(p >= 0.f && p < 1.f, "invalid input, prob should be between 0 and 1 (inclusive)");
    // hash3(seed, stream 77, i) of the synthetic-input generator: base = mix64(seed*K1 + 77*K2), value = mix64(base + i)
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + 77ull * 0xD1B54A32D192ED03ull;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    dropout_kernel<<<stream_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, x, p, z, y);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_adam_step(gnn_ctx_t *ctx, int64_t n, float *p, const float *g, float *m, float *v, float lr, float beta1,
                  float beta2, float eps, float weight_decay, int64_t step) {
    GNN_REQUIRE(ctx && p && g && m && v && n > 0 && step >= 1, "gnn_adam_step: bad argument (step counts from 1)");
    const float bc1 = (float)(1.0 - pow((double)beta1, (double)step)), bc2 = (float)(1.0 - pow((double)beta2, (double)step));
    adam_kernel<<<stream_grid(ctx, n, 256, 1), 256, 0, ctx->stream>>>(n, p, g, m, v, lr, beta1, beta2, eps, weight_decay, bc1,
                                                                     bc2);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_softmax_xent_masked(gnn_ctx_t *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                            const uint8_t *mask, int64_t n_selected, float *loss, float *dZ, int64_t ldd) {
    GNN_REQUIRE(ctx && Z && y && mask && loss, "gnn_softmax_xent_masked: NULL argument");
    GNN_REQUIRE(N > 0 && C > 0 && ldz >= C && n_selected > 0,
                "invalid input, logits must be of rank 2 and targets must be 1D tensor");
    int64_t nblocks = ceil_div(N, 8);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (nblocks > cap) nblocks = cap;
    void *ws = nullptr;
    GNN_TRY(ctx->workspace((size_t)nblocks * 4, &ws));
    const float inv_n = 1.0f / (float)n_selected;
    softmax_xent_masked_kernel<<<(unsigned)nblocks, 256, 0, ctx->stream>>>(N, C, Z, ldz, y, mask, inv_n, dZ, ldd, (float *)ws);
    GNN_LAUNCHED(ctx);
    xent_final_kernel<<<1, 32, 0, ctx->stream>>>((int)nblocks, (const float *)ws, inv_n, loss);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_argmax_correct(gnn_ctx_t *ctx, int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                       const uint8_t *mask, int64_t *count) {
    GNN_REQUIRE(ctx && Z && y && count && N > 0 && C > 0 && ldz >= C, "gnn_argmax_correct: bad argument");
    int64_t nblocks = ceil_div(N, 8);
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (nblocks > cap) nblocks = cap;
    void *ws = nullptr;
    GNN_TRY(ctx->workspace((size_t)nblocks * 4, &ws));
    argmax_correct_kernel<<<(unsigned)nblocks, 256, 0, ctx->stream>>>(N, C, Z, ldz, y, mask, (int32_t *)ws);
    GNN_LAUNCHED(ctx);
    count_final_kernel<<<1, 32, 0, ctx->stream>>>((int)nblocks, (const int32_t *)ws, count);
    GNN_LAUNCHED(ctx);
    return 0;
}

int gnn_binary_f32(gnn_ctx_t *ctx, int op, int64_t rows, int64_t cols, const float *a, int64_t a_rs, int64_t a_cs,
                   const float *b, int64_t b_rs, int64_t b_cs, float *out) {
    GNN_REQUIRE(ctx && a && b && out && rows > 0 && cols > 0 && op >= 0 && op <= GNN_OP_GT, "gnn_binary_f32: bad argument");
    binary_kernel<<<stream_grid(ctx, rows * cols, 256), 256, 0, ctx->stream>>>(op, rows, cols, a, a_rs, a_cs, b, b_rs,
                                                                             b_cs, out);
    GNN_LAUNCHED(ctx);
    return 0;
}
int gnn_unary_f32(gnn_ctx_t *ctx, int op, int64_t n, const float *a, float *out) {
    GNN_REQUIRE(ctx && a && out && n > 0 && op >= 0 && op <= GNN_UOP_NEG, "gnn_unary_f32: bad argument");
    unary_kernel<<<stream_grid(ctx, n, 256), 256, 0, ctx->stream>>>(op, n, a, out);
    GNN_LAUNCHED(ctx);
    return 0;
}
int gnn_where_f32(gnn_ctx_t *ctx, int64_t n, const float *cond, const float *t, const float *f, float *out) {
    GNN_REQUIRE(ctx && cond && t && f && out && n > 0, "gnn_where_f32: bad argument");
    where_kernel<<<stream_grid(ctx, n, 256), 256, 0, ctx->stream>>>(n, cond, t, f, out);
    GNN_LAUNCHED(ctx);
    return 0;
}
int gnn_sum_f32(gnn_ctx_t *ctx, int64_t rows, int64_t cols, const float *a, int dim, float *out) {
    GNN_REQUIRE(ctx && a && out && rows > 0 && cols > 0, "gnn_sum_f32: bad argument");
    if (dim == 0) return colsum(ctx, rows, (int32_t)cols, a, cols, out);
    if (dim == 1) {
        rowsum_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, ctx->stream>>>(rows, cols, a, out);
        GNN_LAUNCHED(ctx);
        return 0;
    }
    // all elements: treat as one long column
    return colsum(ctx, rows * cols, 1, a, 1, out);
}
int gnn_transpose_f32(gnn_ctx_t *ctx, int64_t rows, int64_t cols, const float *a, float *out) {
    GNN_REQUIRE(ctx && a && out && rows > 0 && cols > 0 && a != out, "gnn_transpose_f32: bad argument");
    const int64_t tiles_c = ceil_div(cols, 32), tiles_r = ceil_div(rows, 32);
    dim3 block(32, 8);
    transpose_kernel<<<(unsigned)(tiles_c * tiles_r), block, 0, ctx->stream>>>(rows, cols, tiles_c, a, out);
    GNN_LAUNCHED(ctx);
    return 0;
}
int gnn_gather_cols_f32(gnn_ctx_t *ctx, int64_t rows, int64_t cols, const float *a, const int32_t *idx, float *out) {
    GNN_REQUIRE(ctx && a && idx && out && rows > 0 && cols > 0, "gnn_gather_cols_f32: bad argument");
    gather_cols_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, ctx->stream>>>(rows, cols, a, idx, out);
    GNN_LAUNCHED(ctx);
    return 0;
}

} // extern "C"
