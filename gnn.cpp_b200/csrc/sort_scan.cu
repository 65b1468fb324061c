// sort_scan.cu — device-wide primitives for the structure build (K1/K2): exclusive scan and a stable
// LSD radix sort of 64-bit keys (optional 32-bit payload).  HBM-bound integer work: coalesced tile
// loads, shared-memory histograms, warp match for stable ranking.  Hand-written (no CUB/Thrust).
#include "common.cuh"

namespace gnn {

// ------------------------------------------------------------------------------------------------
// exclusive scan
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// one tile per block: out = exclusive scan inside the tile, sums[block] = tile total
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const uint32_t *__restrict__ in,
                                                                 uint32_t *__restrict__ out, int64_t n,
                                                                 uint32_t *__restrict__ sums,
                                                                 uint32_t *__restrict__ total) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)tid * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t local = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        local += v[i];
    }
    uint32_t incl = warp_incl_scan(local, lane);
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
        uint32_t wi = warp_incl_scan(w, lane);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w; // exclusive
        if (lane == SCAN_THREADS / 32 - 1) {
            if (sums) sums[blockIdx.x] = wi;
            if (total && gridDim.x == 1) *total = wi;
        }
    }
    __syncthreads();
    uint32_t run = warp_sums[warp] + incl - local;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(uint32_t *__restrict__ out, int64_t n,
                                                                const uint32_t *__restrict__ offs) {
    const uint32_t o = offs[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) out[base + i] += o;
}

int exclusive_scan_u32(gnn_ctx *ctx, const uint32_t *in, uint32_t *out, int64_t n, uint32_t *total_d) {
    if (n <= 0) {
        if (total_d) GNN_CHECK_CUDA(cudaMemsetAsync(total_d, 0, 4, ctx->stream));
        return 0;
    }
    const int64_t nb = ceil_div(n, SCAN_TILE);
    if (nb == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, nullptr, total_d);
        GNN_LAUNCHED(ctx);
        return 0;
    }
    uint32_t *sums = nullptr;
    GNN_CHECK_CUDA(cudaMallocAsync((void **)&sums, (size_t)nb * 4, ctx->stream));
    scan_tile_kernel<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, sums, nullptr);
    GNN_LAUNCHED(ctx);
    GNN_TRY(exclusive_scan_u32(ctx, sums, sums, nb, total_d));
    scan_add_kernel<<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(out, n, sums);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaFreeAsync(sums, ctx->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits
// ------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;                     // keys per thread
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per block
constexpr int RS_RADIX = 256;

__global__ void __launch_bounds__(RS_THREADS) rs_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                             uint32_t mask, uint32_t *__restrict__ counts,
                                                             int64_t nb) {
    __shared__ uint32_t hist[RS_RADIX];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; i++) {
        int64_t k = base + (int64_t)i * RS_THREADS + threadIdx.x;
        if (k < n) atomicAdd(&hist[(uint32_t)(keys[k] >> shift) & mask], 1u);
    }
    __syncthreads();
    counts[(int64_t)threadIdx.x * nb + blockIdx.x] = hist[threadIdx.x]; // digit-major for the global scan
}

template <bool HAS_VALS>
__global__ void __launch_bounds__(RS_THREADS)
    rs_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ vals, int64_t n, int shift,
                      uint32_t mask, const uint32_t *__restrict__ offsets, int64_t nb, uint64_t *__restrict__ keys_out,
                      uint32_t *__restrict__ vals_out) {
    // per-warp running digit counts, later turned into per-warp exclusive offsets
    // dynamic shared memory (the payload variant needs 59 KB): staged keys | staged payloads | per-warp histograms | ...
    extern __shared__ __align__(16) uint8_t rs_smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(rs_smem);
    uint32_t *svals = reinterpret_cast<uint32_t *>(rs_smem + RS_TILE * 8);
    uint32_t(*whist)[RS_RADIX] = reinterpret_cast<uint32_t(*)[RS_RADIX]>(rs_smem + RS_TILE * 8 + (HAS_VALS ? RS_TILE * 4 : 0));
    uint32_t *gbase = &whist[RS_WARPS][0], *dstart = gbase + RS_RADIX, *wsum = dstart + RS_RADIX;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int w = 0; w < RS_WARPS; w++) whist[w][tid] = 0;
    gbase[tid] = offsets[(int64_t)tid * nb + blockIdx.x];
    __syncthreads();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * (RS_ITEMS * 32);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int64_t k = wbase + i * 32 + lane;
        const bool valid = k < n;
        key[i] = valid ? keys[k] : ~0ull;
        // invalid lanes use digit RS_RADIX (never matches a real digit) so they do not disturb ranks
        const uint32_t d = valid ? ((uint32_t)(key[i] >> shift) & mask) : (uint32_t)RS_RADIX;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(peers) - 1;
        uint32_t pre = 0;
        if (valid && lane == leader) {
            pre = whist[warp][d];
            whist[warp][d] = pre + __popc(peers);
        }
        pre = __shfl_sync(0xffffffffu, pre, leader);
        rank[i] = pre + __popc(peers & lt_mask);
        __syncwarp();
    }
    __syncthreads();
    // exclusive prefix over warps for digit = tid; dstart[d] = first position of digit d inside the tile's sorted order
    {
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = whist[w][tid];
            whist[w][tid] = run;
            run += c;
        }
        // exclusive scan of the 256 digit totals (one per thread)
        const uint32_t incl = warp_incl_scan(run, lane);
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) before += w < warp ? wsum[w] : 0u;
        dstart[tid] = before + incl - run;
    }
    __syncthreads();
    // stage the tile in digit order in shared memory, then write it out in that order: keys of one digit leave as one
    // contiguous run (a direct scatter writes 8-byte pieces to 256 places and used ~4x the DRAM sectors)
    int64_t tile_n = n - (int64_t)blockIdx.x * RS_TILE;
    if (tile_n > RS_TILE) tile_n = RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_ITEMS; i++) {
        const int64_t k = wbase + i * 32 + lane;
        if (k < n) {
            const uint32_t d = (uint32_t)(key[i] >> shift) & mask;
            const uint32_t j = dstart[d] + whist[warp][d] + rank[i];
            skeys[j] = key[i];
            if (HAS_VALS) svals[j] = vals[k];
        }
    }
    __syncthreads();
    for (int j = tid; j < (int)tile_n; j += RS_THREADS) {
        const uint64_t kk = skeys[j];
        const uint32_t d = (uint32_t)(kk >> shift) & mask;
        const int64_t pos = (int64_t)gbase[d] + ((uint32_t)j - dstart[d]);
        keys_out[pos] = kk;
        if (HAS_VALS) vals_out[pos] = svals[j];
    }
}

static constexpr size_t rs_smem_bytes(bool has_vals) {
    return (size_t)RS_TILE * 8 + (has_vals ? (size_t)RS_TILE * 4 : 0) + (size_t)RS_WARPS * RS_RADIX * 4 + 2 * RS_RADIX * 4 + RS_WARPS * 4;
}

int radix_sort_u64(gnn_ctx *ctx, uint64_t *keys, uint32_t *vals, int64_t n, int bit_lo, int bit_hi) {
    if (n <= 1 || bit_hi <= bit_lo) return 0;
    static uint64_t attr_set = 0; // per device: the payload variant needs more than the default 48 KB of shared memory
    if (!(attr_set >> (ctx->device & 63) & 1)) {
        GNN_CHECK_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true)));
        GNN_CHECK_CUDA(cudaFuncSetAttribute(rs_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(false)));
        attr_set |= 1ull << (ctx->device & 63);
    }
    GNN_REQUIRE(n < (int64_t)0xFFFFFFFFll, "radix_sort_u64: n=%lld exceeds 32-bit offsets", (long long)n);
    const int64_t nb = ceil_div(n, RS_TILE);
    const size_t counts_bytes = (size_t)round_up(nb * RS_RADIX * 4, 256);
    const size_t keys_bytes = (size_t)round_up(n * 8, 256);
    const size_t vals_bytes = vals ? (size_t)round_up(n * 4, 256) : 0;
    char *ws = nullptr;
    GNN_TRY(ctx->workspace(counts_bytes + keys_bytes + vals_bytes, (void **)&ws));
    uint32_t *counts = (uint32_t *)ws;
    uint64_t *kalt = (uint64_t *)(ws + counts_bytes);
    uint32_t *valt = vals ? (uint32_t *)(ws + counts_bytes + keys_bytes) : nullptr;
    uint64_t *kin = keys, *kout = kalt;
    uint32_t *vin = vals, *vout = valt;
    for (int shift = bit_lo; shift < bit_hi; shift += 8) {
        const int bits = bit_hi - shift < 8 ? bit_hi - shift : 8;
        const uint32_t mask = (1u << bits) - 1u;
        rs_hist_kernel<<<(unsigned)nb, RS_THREADS, 0, ctx->stream>>>(kin, n, shift, mask, counts, nb);
        GNN_LAUNCHED(ctx);
        GNN_TRY(exclusive_scan_u32(ctx, counts, counts, nb * RS_RADIX, nullptr));
        if (vals)
            rs_scatter_kernel<true><<<(unsigned)nb, RS_THREADS, rs_smem_bytes(true), ctx->stream>>>(kin, vin, n, shift, mask, counts,
                                                                                                 nb, kout, vout);
        else
            rs_scatter_kernel<false><<<(unsigned)nb, RS_THREADS, rs_smem_bytes(false), ctx->stream>>>(kin, nullptr, n, shift, mask,
                                                                                                   counts, nb, kout, nullptr);
        GNN_LAUNCHED(ctx);
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tv = vin; vin = vout; vout = tv;
    }
    if (kin != keys) {
        GNN_CHECK_CUDA(cudaMemcpyAsync(keys, kin, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (vals) GNN_CHECK_CUDA(cudaMemcpyAsync(vals, vin, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return 0;
}

} // namespace gnn
