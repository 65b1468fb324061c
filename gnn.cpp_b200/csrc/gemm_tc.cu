// gemm_tc.cu — tensor-core (tcgen05, 3xTF32) variants of the dense feature transforms K6.
//
// Every FP32 operand a is split as a = hi + lo; the product is accumulated in FP32 (TMEM) as
// lo_a*hi_b + hi_a*lo_b + hi_a*hi_b  (error ~2^-20, inside the 1e-5 contract; the dropped lo*lo term is ~2^-21).
// The small operand (weights) is pre-split once per call with hi = rna_tf32(b), lo = rna_tf32(b - hi).  The big,
// streamed operand is split ON CHIP: TMA lands the raw FP32 tile in shared memory (128-byte swizzle) and converter
// warps write lo next to it.  Default (TRUNC): the tensor core only reads the upper 19 bits of a TF32 operand, so the
// raw tile IS hi = trunc(a) and only lo = rna_tf32(a - trunc(a)) is written (3 instructions per element); TRUNC = false is the
// round-to-nearest split (hi rewritten in place; two emulated cvt.rna per element).  One thread issues the tcgen05.mma
// triple.
//
//   rows kernel (NT, NN):  C[M,N] = A[M,K] * Bt[N,K]^T   A streamed by 128-row tiles, persistent CTAs,
//                          two TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1,
//                          fused bias / ReLU / mask epilogue with coalesced row stores.
//   tn kernel:             C[K1,K2] = A[M,K1]^T * B[M,K2] reduction over the (millions of) node rows: both operands
//                          are MN-major straight out of TMA, every CTA owns a contiguous node range and writes one
//                          partial; a fixed-order pass sums the partials (deterministic, no atomics).
//
// Entry points return -1 when a shape/alignment is not supported so gemm.cu falls back to the FP32 FMA kernel.
#include <cuda.h>

#include "common.cuh"

namespace gnn {
namespace tc {

// ------------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the device
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *tm, uint32_t bar, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 :
                 : "r"(dst), "l"((uint64_t)tm), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// smem -> global tile store through the async proxy (no LSU instructions, completion tracked by bulk groups)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, uint32_t src, int32_t c0, int32_t c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :
                 : "l"((uint64_t)tm), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// at most one committed store group may still be reading its shared-memory source
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tm) : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of one TPC run ONE tcgen05.mma of 256 rows; each
// CTA stages its own 128 rows of A and only HALF of B in shared memory (the pair's tensor cores exchange the halves), so
// the per-SM shared-memory traffic of the B operand — TMA fill and MMA reads — is halved.  CTA rank 0 (the leader)
// issues the MMAs and owns the barriers the other CTA signals remotely.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ uint32_t ld_shared_cluster_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// arrive on a barrier that may live in the other CTA of the pair (release at cluster scope)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// the same with the default (.release.cta) semantics — what CUTLASS' ClusterBarrier::arrive(cta_id) issues; no
// MEMBAR.GPU in front of the arrive (see pair_publish)
__device__ __forceinline__ void mbar_arrive_cluster_light(uint32_t cluster_bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// wait on a local barrier whose arrivals come from both CTAs of the pair (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// fence between the generic and the async proxy for every state space (pair mode: the tile this CTA rewrote is consumed
// by an MMA the OTHER CTA issues)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// A warp of either CTA publishes "my part of this stage is in shared memory, visible to the tensor core" on the
// leader's barrier, once per pipeline stage.  Default: proxy fence on the CTA's own shared memory (where the tile lives
// and where the tensor core reads it) + a default-scope remote arrive — the sequence CUTLASS' 2-SM kernels use for
// operand tiles written by threads.  The formally stronger form (full-space proxy fence + release at cluster scope) puts
// two MEMBAR.ALL.GPU on the path of every stage and made the pair kernels 1.5x SLOWER than the single-CTA ones
// (profiles/r2b_gemm_pair.md); it is kept behind GNN_GEMM_DEBUG bit 8 for A/B runs.
__device__ __forceinline__ void pair_publish(uint32_t leader_bar, int lane, bool heavy) {
    if (!heavy) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_light(leader_bar);
    } else {
        fence_proxy_async_all();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_bar);
    }
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// The pair kernels also allocate per CTA with these cta_group::1 forms: the same amount from an empty TMEM (one CTA per
// SM) lands on the same columns in both CTAs, which is what one cta_group::2 MMA needs; the kernels check it and trap.
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :
                 : "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// arrives on the mbarrier when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// pair forms: one MMA over both CTAs' operands (issued by the leader); the commit arrives on the barrier at the same
// shared-memory offset in BOTH CTAs
__device__ __forceinline__ void mma_tf32_pair(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    const uint32_t z = 0;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                 :
                 : "r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
template <int NCTA> __device__ __forceinline__ void mma_issue(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if constexpr (NCTA == 2) mma_tf32_pair(d, da, db, idesc, acc);
    else mma_tf32(d, da, db, idesc, acc);
}
template <int NCTA> __device__ __forceinline__ void mma_done(uint32_t bar) {
    if constexpr (NCTA == 2) mma_commit_pair(bar);
    else mma_commit(bar);
}
// 32 lanes x 16 consecutive 32-bit columns: thread i of the warp gets lane (base+i), columns c..c+15
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t *r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t rna_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
// hi (rounded to TF32) and lo (the rounded remainder) of one 16-byte vector already in registers
__device__ __forceinline__ void split4_from(const float4 v, float4 *hi_ptr, float4 *lo_ptr) {
    float4 h, l;
    h.x = __uint_as_float(rna_tf32(v.x)); l.x = __uint_as_float(rna_tf32(v.x - h.x));
    h.y = __uint_as_float(rna_tf32(v.y)); l.y = __uint_as_float(rna_tf32(v.y - h.y));
    h.z = __uint_as_float(rna_tf32(v.z)); l.z = __uint_as_float(rna_tf32(v.z - h.z));
    h.w = __uint_as_float(rna_tf32(v.w)); l.w = __uint_as_float(rna_tf32(v.w - h.w));
    *hi_ptr = h;
    *lo_ptr = l;
}

// Truncating split: the tensor core reads only the upper 19 bits of a TF32 operand, so the raw FP32 tile as TMA landed
// it already IS the hi operand (hi = trunc(v): the low 13 mantissa bits are ignored); only lo has to be written.
// d = v - trunc(v) is exact in FP32; adding 0x1000 to its bit pattern makes the hardware's truncation of lo a
// round-to-nearest (ties away), so lo = rna_tf32(d) for one integer add.  3 instructions per element instead of two
// emulated cvt.rna (~5 SASS instructions each) + subtract, and no rewrite of the hi tile.
// |v - (hi + lo)| <= 2^-21 |v|, zero-mean (round-to-nearest split: 2^-22).  (d = 0 -> 0x1000, which the tensor core
// reads as 0; a non-finite v leaves hi non-finite, which is what propagates.)
__device__ __forceinline__ float lo_of(float v) {
    const float d = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    return __uint_as_float(__float_as_uint(d) + 0x1000u);
}
__device__ __forceinline__ void split4_trunc(const float4 v, float4 *lo_ptr) {
    *lo_ptr = make_float4(lo_of(v.x), lo_of(v.y), lo_of(v.z), lo_of(v.w));
}

// shared-memory matrix descriptor, descriptor version 1 (sm_100).  layout 2 = 128-byte swizzle of 16-byte chunks
// (TMA SWIZZLE_128B), layout 1 = 128-byte swizzle of 32-byte chunks (TMA SWIZZLE_128B_ATOM_32B) — the only layout
// the tensor core accepts for MN-major 32-bit operands.
constexpr uint64_t LAYOUT_SW128 = 2, LAYOUT_SW128_BASE32B = 1, LAYOUT_SW64 = 4;
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint64_t layout = LAYOUT_SW128) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    d |= layout << 61;
    return d;
}
// instruction descriptor: D=F32, A=B=TF32, M x N tile, operand majors (0 = K-major, 1 = MN-major)
__host__ __device__ constexpr uint32_t instr_desc(uint32_t M, uint32_t N, uint32_t a_mn, uint32_t b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn << 15) | (b_mn << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// pipeline position: stage index and phase parity advanced incrementally (a runtime `it % stages` is an integer division
// on the per-stage path of every warp role — ~25 instructions in the ncu source view)
struct StagePos {
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ void next(uint32_t stages) {
        if (++s == stages) {
            s = 0;
            ph ^= 1u;
        }
    }
};

// ------------------------------------------------------------------------------------------------ rows kernel
constexpr int ROWS_THREADS = 320; // warp 0 TMA, warp 1 MMA, warps 2-5 converters, warps 6-9 epilogue
constexpr int TILE_M = 128;
constexpr int BK = 32;                      // floats per k-block = one 128-byte swizzle row
constexpr uint32_t A_TILE_BYTES = TILE_M * BK * 4; // 16 KB
constexpr uint32_t CTRL_BYTES = 2048;            // [0,1024) mbarriers + TMEM slot, [1024,2048) bias copy (256 floats)
constexpr uint32_t STAGING_BYTES = 4 * 32 * 128; // 4 epilogue warps x 32 rows x 32 columns (x2 buffers, x2 with a mask, in TMA-epilogue mode)

struct RowsArgs {
    int64_t M;
    int32_t N, Npad, kblocks, num_tiles, stages, bk; // bk = floats per k-block: 32 (128-byte swizzle) or 16 (64-byte)
    uint32_t tmem_cols, acc_stride, stage_bytes, a_bytes, b_bytes;
    uint32_t epi_bytes;  // shared memory of the epilogue: staging (+ mask tiles)
    int tma_epi;         // 1: output tiles leave (and mask tiles arrive) through TMA; 0: per-lane global stores/loads
    uint32_t debug;      // timing experiments only (GNN_GEMM_DEBUG): 1 skip global stores, 2 skip the hi/lo split, 4 skip mask loads
    uint32_t b_resident; // bytes of the whole split weight matrix kept in shared memory for the kernel's lifetime (0 = streamed per k-block)
    float *C;
    int64_t ldc;
    const float *bias;
    int relu;
    const float *mask;
    int64_t ldm;
    // bias gradient fused into the epilogue (TMA epilogue only): per epilogue warp the column sums of the tiles it wrote,
    // [gridDim.x * 4][256]; a fixed-order pass adds the partials (rows_colsum_final_kernel).  NULL = not wanted.
    float *colsum_part;
};

// Column sums over the 32 rows a warp holds (lane = row, x[j] = column j): a butterfly that halves the columns a lane
// keeps at every step — after 5 steps lane j holds, in x[0], the sum of column j over all 32 lanes.  31 shuffles, fixed
// order (deterministic).
__device__ __forceinline__ float warp_column_sums(float (&x)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool upper = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; i++) {
            const float send = upper ? x[i] : x[i + s];
            const float keep = upper ? x[i + s] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return x[0];
}

// NCTA = 1: one CTA per 128-row tile.  NCTA = 2: a CTA pair (cluster of 2, launched with the cluster attribute) per
// 256-row tile — every CTA runs the same roles on its own 128 rows and its own half of the weights' output columns;
// only the leader's MMA thread issues (cta_group::2, M = 256).  Barriers signalled by threads of both CTAs (converted
// tile ready, accumulator drained) live in the leader and count both CTAs' warps; barriers signalled by the tensor core
// (stage free, accumulator full) are local in each CTA and receive a multicast commit.
template <int NCTA, bool TRUNC>
__global__ void __launch_bounds__(ROWS_THREADS, 1)
    tc_rows_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmM, const RowsArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw_u32);

    const uint32_t bar_full = base, bar_conv = base + 64, bar_empty = base + 128;
    const uint32_t bar_tfull = base + 192, bar_tempty = base + 208, tmem_slot = base + 224, bar_bres = base + 232;
    const uint32_t bar_mask = base + 256; // 4 epilogue warps x 2 mask buffers
    const uint32_t staging = base + CTRL_BYTES;
    const uint32_t bres = staging + a.epi_bytes;          // resident weights (when a.b_resident != 0)
    const uint32_t stage0 = bres + a.b_resident;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = NCTA == 2 ? cluster_ctarank() : 0u;
    // the leader's copies of the barriers that threads of both CTAs arrive on
    const uint32_t bar_conv_l = NCTA == 2 ? mapa_shared(bar_conv, 0) : bar_conv;
    const uint32_t bar_tempty_l = NCTA == 2 ? mapa_shared(bar_tempty, 0) : bar_tempty;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        mbar_init(bar_bres, 1);
        for (int i = 0; i < 8; i++) mbar_init(bar_mask + 8 * i, 1);
        if (a.tma_epi) {
            tma_prefetch_desc(&tmC);
            if (a.mask) tma_prefetch_desc(&tmM);
        }
        for (int s = 0; s < a.stages; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 4 * NCTA);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4 * NCTA);
        }
        fence_barrier_init();
    }
    if (warp >= 6) { // bias copy for the TMA epilogue (zero when there is no bias / beyond N)
        float *bias_s = reinterpret_cast<float *>(gbase + 1024);
        for (int i = threadIdx.x - 192; i < 256; i += 128) bias_s[i] = (a.bias && i < a.N) ? __ldg(a.bias + i) : 0.f;
    }
    if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    if constexpr (NCTA == 2) cluster_sync_all(); // also: the other CTA's barriers exist before anything signals them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gbase + 224);
    if constexpr (NCTA == 2) {
        // one MMA addresses the accumulator at the same TMEM columns in both CTAs: each CTA allocates the same amount
        // from an empty TMEM (1 CTA per SM), so the bases agree — anything else must fail loudly
        if (threadIdx.x == 0 && ld_shared_cluster_u32(mapa_shared(tmem_slot, cta_rank ^ 1u)) != tmem_base) __trap();
    }

    // tiles are NCTA * 128 rows; this CTA's rows of tile t start at row_of(t)
    const int32_t first_tile = blockIdx.x / NCTA, tile_step = gridDim.x / NCTA;
    auto row_of = [&](int32_t tile) { return (tile * NCTA + (int32_t)cta_rank) * TILE_M; };
    const int32_t b_row0 = (int32_t)cta_rank * (a.Npad / NCTA); // this CTA's rows of the (hi | lo) weight halves

    if (warp == 0) {
        if (lane == 0) {
            StagePos pos;
            if (a.b_resident) { // the whole split weight matrix (this CTA's output columns), once: [k-block][hi | lo]
                mbar_arrive_expect_tx(bar_bres, a.b_resident);
                for (int32_t kb = 0; kb < a.kblocks; kb++) {
                    tma_load_2d(bres + kb * 2 * a.b_bytes, &tmB, bar_bres, kb * a.bk, b_row0);
                    tma_load_2d(bres + kb * 2 * a.b_bytes + a.b_bytes, &tmB, bar_bres, kb * a.bk, a.Npad + b_row0);
                }
            }
            const uint32_t tx = a.a_bytes + (a.b_resident ? 0 : 2 * a.b_bytes);
            for (int32_t tile = first_tile; tile < a.num_tiles; tile += tile_step) {
                for (int32_t kb = 0; kb < a.kblocks; kb++, pos.next(a.stages)) {
                    const uint32_t s = pos.s, ph = pos.ph;
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    const uint32_t st = stage0 + s * a.stage_bytes;
                    mbar_arrive_expect_tx(bar_full + 8 * s, tx);
                    tma_load_2d(st, &tmA, bar_full + 8 * s, kb * a.bk, row_of(tile));
                    if (!a.b_resident) {
                        tma_load_2d(st + 2 * a.a_bytes, &tmB, bar_full + 8 * s, kb * a.bk, b_row0);
                        tma_load_2d(st + 2 * a.a_bytes + a.b_bytes, &tmB, bar_full + 8 * s, kb * a.bk, a.Npad + b_row0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            const uint32_t idesc = instr_desc(TILE_M * NCTA, (uint32_t)a.Npad, 0, 0);
            uint32_t tile_it = 0;
            StagePos pos;
            if (NCTA == 1 && a.b_resident) mbar_wait(bar_bres, 0);
            for (int32_t tile = first_tile; tile < a.num_tiles; tile += tile_step, tile_it++) {
                const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
                if (NCTA == 2 && (a.debug & 8)) mbar_wait_cluster(bar_tempty + 8 * acc, acc_ph ^ 1);
                else mbar_wait(bar_tempty + 8 * acc, acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * a.acc_stride;
                for (int32_t kb = 0; kb < a.kblocks; kb++, pos.next(a.stages)) {
                    const uint32_t s = pos.s, ph = pos.ph;
                    if constexpr (NCTA == 2) {
                        // the converter warps of BOTH CTAs arrive here, each after it saw its own CTA's stage (A tile
                        // and weight half) land: one wait covers all four operand tiles of the pair
                        if (a.debug & 8) mbar_wait_cluster(bar_conv + 8 * s, ph);
                        else mbar_wait(bar_conv + 8 * s, ph);
                    } else {
                        mbar_wait(bar_full + 8 * s, ph);
                        mbar_wait(bar_conv + 8 * s, ph);
                    }
                    tc_fence_after();
                    const uint32_t st = stage0 + s * a.stage_bytes;
                    // K-major operands: 8-row groups are sbo bytes apart (8 rows x one swizzle row of bk floats)
                    const uint32_t sbo = 32u * a.bk;
                    const uint64_t lay = a.bk == 32 ? LAYOUT_SW128 : LAYOUT_SW64;
                    const uint64_t a_hi = smem_desc(st, 16, sbo, lay), a_lo = smem_desc(st + a.a_bytes, 16, sbo, lay);
                    const uint32_t bst = a.b_resident ? bres + kb * 2 * a.b_bytes : st + 2 * a.a_bytes;
                    const uint64_t b_hi = smem_desc(bst, 16, sbo, lay);
                    const uint64_t b_lo = smem_desc(bst + a.b_bytes, 16, sbo, lay);
                    for (uint32_t k = 0; k < (uint32_t)a.bk / 8; k++) {
                        const uint64_t adv = (uint64_t)(k * 32 >> 4); // 8 tf32 = 32 bytes along the swizzled row
                        mma_issue<NCTA>(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb | k) != 0);
                        mma_issue<NCTA>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                        mma_issue<NCTA>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    mma_done<NCTA>(bar_empty + 8 * s);
                }
                mma_done<NCTA>(bar_tfull + 8 * acc);
            }
        }
    } else if (warp < 6) {
        const int t = threadIdx.x - 64; // 0..127
        StagePos pos;
        if (NCTA == 2 && a.b_resident) mbar_wait(bar_bres, 0); // pair mode: "converted" also vouches for the weights
        for (int32_t tile = first_tile; tile < a.num_tiles; tile += tile_step) {
            for (int32_t kb = 0; kb < a.kblocks; kb++, pos.next(a.stages)) {
                const uint32_t s = pos.s, ph = pos.ph;
                mbar_wait(bar_full + 8 * s, ph);
                uint8_t *st = gbase + CTRL_BYTES + a.epi_bytes + a.b_resident + (size_t)s * a.stage_bytes;
                const int nit = (a.debug & 2) ? 0 : a.bk / 4; // 128 x bk floats = 32 bk float4 over 128 threads
                for (int i0 = 0; i0 < nit; i0 += 4) { // nit is 4 or 8; four loads in flight, then the conversions and stores
                    float4 raw[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) raw[j] = reinterpret_cast<const float4 *>(st)[(i0 + j) * 128 + t];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float4 *hp = reinterpret_cast<float4 *>(st) + ((i0 + j) * 128 + t);
                        if constexpr (TRUNC) split4_trunc(raw[j], hp + a.a_bytes / 16);
                        else split4_from(raw[j], hp, hp + a.a_bytes / 16);
                    }
                }
                if constexpr (NCTA == 2) {
                    pair_publish(bar_conv_l + 8 * s, lane, a.debug & 8);
                } else {
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_conv + 8 * s);
                }
            }
        }
    } else if (a.tma_epi) {
        // Epilogue through the async proxy: a warp owns the 32 accumulator rows of its TMEM lane quadrant; per chunk
        // of 32 columns every lane (= one row) applies bias / ReLU / mask in registers, writes the row into a
        // 128-byte-swizzled staging tile, and one lane hands the tile to a TMA store.  Mask tiles are fetched two
        // chunks ahead by TMA loads.  No global load/store instruction is issued here, so the LSU queue the
        // converter warps share stays free (per-lane STG/LDG made the epilogue additive to the MMA time).
        const int q = warp & 3;
        const uint32_t stg0 = staging + q * 8192;                          // two 4 KB staging tiles
        const uint32_t msk0 = staging + 32768 + q * 8192;                  // two 4 KB mask tiles (only with a mask)
        uint8_t *g_stg = gbase + CTRL_BYTES + q * 8192, *g_msk = gbase + CTRL_BYTES + 32768 + q * 8192;
        const float *bias_s = reinterpret_cast<const float *>(gbase + 1024);
        const uint32_t mbar = bar_mask + 16 * q;
        const int32_t nch = (a.Npad + 31) / 32;                            // chunks per tile
        const int32_t my_tiles = first_tile < a.num_tiles ? (a.num_tiles - first_tile + tile_step - 1) / tile_step : 0;
        const int64_t total_ch = (int64_t)my_tiles * nch;
        auto issue_mask = [&](int64_t ci) { // lane 0: TMA load of the mask tile of chunk ci into buffer ci & 1
            const int32_t tile = first_tile + (int32_t)(ci / nch) * tile_step;
            const int32_t c0 = (int32_t)(ci % nch) * 32;
            mbar_arrive_expect_tx(mbar + 8 * (ci & 1), 4096);
            tma_load_2d(msk0 + (uint32_t)(ci & 1) * 4096, &tmM, mbar + 8 * (ci & 1), c0, row_of(tile) + q * 32);
        };
        if (a.mask && lane == 0)
            for (int64_t ci = 0; ci < 2 && ci < total_ch; ci++) issue_mask(ci);
        float csum[8]; // fused bias gradient: lane j, slot k = running sum of output column 32 k + j over this warp's rows
#pragma unroll
        for (int k = 0; k < 8; k++) csum[k] = 0.f;
        int64_t ci = 0;
        uint32_t tile_it = 0;
        for (int32_t tile = first_tile; tile < a.num_tiles; tile += tile_step, tile_it++) {
            const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
            mbar_wait(bar_tfull + 8 * acc, acc_ph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * a.acc_stride;
            const int32_t row0 = row_of(tile) + q * 32;
            for (int32_t c0 = 0; c0 < a.Npad; c0 += 32, ci++) {
                const uint32_t b = (uint32_t)(ci & 1);
                uint32_t r[32];
                const bool two = c0 + 16 < a.Npad;
                tmem_ld16(t_row + c0, r);
                if (two) tmem_ld16(t_row + c0 + 16, r + 16);
                else {
#pragma unroll
                    for (int i = 16; i < 32; i++) r[i] = 0u;
                }
                tmem_ld_wait();
                if (c0 + 32 >= a.Npad) { // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        // nothing but "my tcgen05.ld of this accumulator are complete" is published: no memory fence
                        if constexpr (NCTA == 2) mbar_arrive_cluster_light(bar_tempty_l + 8 * acc);
                        else mbar_arrive(bar_tempty + 8 * acc);
                    }
                }
                // the TMA store issued two chunks ago must have finished reading staging tile b
                if (lane == 0) bulk_wait_read1();
                if (a.mask) mbar_wait(mbar + 8 * b, (uint32_t)(ci >> 1) & 1);
                __syncwarp();
                uint8_t *srow = g_stg + b * 4096 + lane * 128;
                const uint8_t *mrow = g_msk + b * 4096 + lane * 128;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    float4 v = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]),
                                           __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
                    const float4 bv = *reinterpret_cast<const float4 *>(bias_s + c0 + 4 * c);
                    v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                    if (a.relu) {
                        v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
                        v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
                    }
                    const int sw = (c ^ (lane & 7)) << 4;
                    if (a.mask) {
                        const float4 mk = *reinterpret_cast<const float4 *>(mrow + sw);
                        v.x = mk.x > 0.f ? v.x : 0.f; v.y = mk.y > 0.f ? v.y : 0.f;
                        v.z = mk.z > 0.f ? v.z : 0.f; v.w = mk.w > 0.f ? v.w : 0.f;
                    }
                    *reinterpret_cast<float4 *>(srow + sw) = v;
                    if (a.colsum_part) { // keep what was written (rows past M contribute nothing)
                        r[4 * c] = __float_as_uint(v.x); r[4 * c + 1] = __float_as_uint(v.y);
                        r[4 * c + 2] = __float_as_uint(v.z); r[4 * c + 3] = __float_as_uint(v.w);
                    }
                }
                fence_proxy_async(); // staging writes (generic proxy) -> visible to the TMA store (async proxy)
                __syncwarp();
                if (lane == 0) {
                    if (!(a.debug & 1)) tma_store_2d(&tmC, stg0 + b * 4096, c0, row0);
                    bulk_commit();
                    if (a.mask && ci + 2 < total_ch) issue_mask(ci + 2); // mask tile b has been consumed by every lane
                }
                if (a.colsum_part) {
                    const bool live = (int64_t)row0 + lane < a.M;
                    float x[32];
#pragma unroll
                    for (int i = 0; i < 32; i++) x[i] = live ? __uint_as_float(r[i]) : 0.f;
                    const float colv = warp_column_sums(x, lane);
                    const int ch = c0 >> 5;
#pragma unroll
                    for (int k = 0; k < 8; k++) csum[k] += k == ch ? colv : 0.f;
                }
            }
        }
        if (a.colsum_part) {
            float *dst = a.colsum_part + ((size_t)blockIdx.x * 4 + q) * 256;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if (k * 32 < a.Npad) dst[k * 32 + lane] = csum[k];
        }
        if (lane == 0) bulk_wait_all();
    } else {
        const int q = warp & 3; // TMEM lane quadrant this warp may read
        uint8_t *stg = gbase + CTRL_BYTES + q * 4096;
        const int rsub = lane >> 3, j = lane & 7;
        uint32_t tile_it = 0;
        for (int32_t tile = first_tile; tile < a.num_tiles; tile += tile_step, tile_it++) {
            const uint32_t acc = tile_it & 1, acc_ph = (tile_it >> 1) & 1;
            mbar_wait(bar_tfull + 8 * acc, acc_ph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * a.acc_stride;
            const int64_t row0 = (int64_t)row_of(tile) + q * 32;
            for (int32_t c0 = 0; c0 < a.Npad; c0 += 32) {
                uint32_t r[32];
                const bool two = c0 + 16 < a.Npad;
                tmem_ld16(t_row + c0, r);
                if (two) tmem_ld16(t_row + c0 + 16, r + 16);
                tmem_ld_wait();
                if (c0 + 32 >= a.Npad) { // accumulator fully read: hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        // nothing but "my tcgen05.ld of this accumulator are complete" is published: no memory fence
                        if constexpr (NCTA == 2) mbar_arrive_cluster_light(bar_tempty_l + 8 * acc);
                        else mbar_arrive(bar_tempty + 8 * acc);
                    }
                }
                // phase 1: thread = row, scatter its 32 columns into the warp's staging rows (xor-swizzled chunks)
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    if (c < 4 || two) {
                        float4 v = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]),
                                               __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
                        *reinterpret_cast<float4 *>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = v;
                    }
                }
                __syncwarp();
                // phase 2: 8 lanes cover one 128-byte row segment -> coalesced global stores, fused epilogue
                const int32_t col = c0 + 4 * j;
                if (col < a.N) {
                    const bool full = col + 3 < a.N;
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (a.bias) { // scalar loads: the bias vector sits anywhere in the parameter slab
                        bv.x = __ldg(a.bias + col);
                        if (col + 1 < a.N) bv.y = __ldg(a.bias + col + 1);
                        if (col + 2 < a.N) bv.z = __ldg(a.bias + col + 2);
                        if (col + 3 < a.N) bv.w = __ldg(a.bias + col + 3);
                    }
                    // all mask vectors of this chunk are fetched before the first store so the loads overlap
                    float4 mk[8];
                    if (a.mask && full && !(a.debug & 4)) {
#pragma unroll
                        for (int itr = 0; itr < 8; itr++) {
                            const int64_t grow = row0 + itr * 4 + rsub;
                            mk[itr] = grow < a.M ? __ldg(reinterpret_cast<const float4 *>(a.mask + grow * a.ldm + col))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int itr = 0; itr < 8; itr++) {
                        const int rr = itr * 4 + rsub;
                        const int64_t grow = row0 + rr;
                        if (grow >= a.M || (a.debug & 1)) continue;
                        float4 v = *reinterpret_cast<const float4 *>(stg + rr * 128 + ((j ^ (rr & 7)) << 4));
                        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
                        if (a.relu) {
                            v.x = v.x > 0.f ? v.x : 0.f; v.y = v.y > 0.f ? v.y : 0.f;
                            v.z = v.z > 0.f ? v.z : 0.f; v.w = v.w > 0.f ? v.w : 0.f;
                        }
                        float *dst = a.C + grow * a.ldc + col;
                        if (full) {
                            if (a.mask) {
                                v.x = mk[itr].x > 0.f ? v.x : 0.f; v.y = mk[itr].y > 0.f ? v.y : 0.f;
                                v.z = mk[itr].z > 0.f ? v.z : 0.f; v.w = mk[itr].w > 0.f ? v.w : 0.f;
                            }
                            *reinterpret_cast<float4 *>(dst) = v;
                        } else {
                            const float vv[3] = {v.x, v.y, v.z};
                            for (int e = 0; e < 3 && col + e < a.N; e++) {
                                float o = vv[e];
                                if (a.mask) o = a.mask[grow * a.ldm + col + e] > 0.f ? o : 0.f;
                                dst[e] = o;
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncwarp(); // the single-lane roles rejoin their warps before the (warp-aligned) barrier
    if constexpr (NCTA == 2) cluster_sync_all(); // neither CTA's shared memory / TMEM goes away under the other's feet
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ tn kernel
// The tensor core adds into its FP32 accumulator with truncation, so a reduction over millions of node rows cannot
// stay in one TMEM accumulator (measured: error grows linearly, ~2e-6 per 96 accumulations).  The node range of a
// CTA is therefore cut into chunks of TN_CHUNK stages; chunks alternate between two TMEM accumulators, and while
// the MMAs of chunk c+1 run, the converter warps drain chunk c into per-thread FP32 register sums (round to
// nearest).  grid = (node splits, 128-row halves of the output).
constexpr int TN_THREADS = 320; // warp 0 TMA, warp 1 MMA, warps 2-9 converters + accumulator drain
constexpr int TN_BK = 16;       // node rows per stage (two K=8 MMA steps)
constexpr int TN_CHUNK = 16;    // stages per TMEM accumulation chunk (256 node rows, 96 accumulating MMAs)
constexpr uint32_t BOX_BYTES = TN_BK * 128; // one TMA box: 16 node rows x 32 floats

struct TnArgs {
    int64_t M, nodes_per_cta;
    int32_t K1, nbB, N, stages; // nbB = B boxes (32 columns each) staged by ONE CTA; N = MMA width (all of B's columns)
    uint32_t tmem_cols, stage_bytes, hi_bytes;
    int64_t part_stride; // floats per node split: halves * 128 * N
    float *partial;      // [splits][halves*128][N]
    uint32_t debug;      // GNN_GEMM_DEBUG bit 8: cluster-scope release for the pair synchronisation (see pair_publish)
};

// NCTA = 2 (K1 > 128): the two 128-row halves of the output are a CTA pair (cluster (2,1,1)) on ONE cta_group::2 MMA of
// M = 256 — each CTA stages and splits its half of A's columns and only HALF of B's columns (NCTA = 1 stages all of B in
// both halves' CTAs), and keeps its 128 output rows in its own TMEM.
template <int NCTA, bool TRUNC>
__global__ void __launch_bounds__(TN_THREADS, 1)
    tc_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw_u32);

    const uint32_t bar_full = base, bar_conv = base + 64, bar_empty = base + 128;
    const uint32_t bar_tfull = base + 192, bar_tempty = base + 208, tmem_slot = base + 224;
    const uint32_t stage0 = base + CTRL_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // NCTA = 1: grid (node splits, halves).  NCTA = 2: grid (2 * node splits), cluster (2,1,1): rank in the pair = half
    const int32_t half = NCTA == 2 ? (int32_t)(blockIdx.x & 1) : (int32_t)blockIdx.y;
    const int32_t split = NCTA == 2 ? (int32_t)(blockIdx.x >> 1) : (int32_t)blockIdx.x;
    const uint32_t bar_conv_l = NCTA == 2 ? mapa_shared(bar_conv, 0) : bar_conv;
    const uint32_t bar_tempty_l = NCTA == 2 ? mapa_shared(bar_tempty, 0) : bar_tempty;
    const int32_t colsA = min(128, a.K1 - half * 128);  // output rows of this half
    const int32_t nbA = (colsA + 31) / 32;              // real A boxes (of the 4 the MMA reads)

    // operand rows/columns beyond K1/K2 only feed ignored accumulator entries; zero them once anyway
    {
        float4 *z = reinterpret_cast<float4 *>(gbase + CTRL_BYTES);
        const uint32_t n16 = (uint32_t)a.stages * a.stage_bytes / 16;
        for (uint32_t i = threadIdx.x; i < n16; i += TN_THREADS) z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (NCTA == 2) fence_proxy_async_all();
        else fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < a.stages; s++) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_conv + 8 * s, 8 * NCTA);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; i++) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 8 * NCTA);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    if constexpr (NCTA == 2) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gbase + 224);
    if constexpr (NCTA == 2) { // see tc_rows_kernel: the pair's accumulators must sit at the same TMEM columns
        if (threadIdx.x == 0 && ld_shared_cluster_u32(mapa_shared(tmem_slot, (uint32_t)half ^ 1u)) != tmem_base) __trap();
    }
    const int32_t b_box0 = NCTA == 2 ? half * a.nbB : 0; // first of B's column boxes this CTA stages

    const int64_t n_begin = (int64_t)split * a.nodes_per_cta;
    const int64_t n_end = min(a.M, n_begin + a.nodes_per_cta);
    const int32_t iters = (int32_t)((n_end - n_begin + TN_BK - 1) / TN_BK);
    const int32_t nchunks = (iters + TN_CHUNK - 1) / TN_CHUNK;
    constexpr uint32_t A_HI_BYTES = 4 * BOX_BYTES;

    if (warp == 0) {
        if (lane == 0) {
            const uint32_t tx = (uint32_t)(nbA + a.nbB) * BOX_BYTES;
            StagePos pos;
            for (int32_t it = 0; it < iters; it++, pos.next(a.stages)) {
                const uint32_t s = pos.s, ph = pos.ph;
                mbar_wait(bar_empty + 8 * s, ph ^ 1);
                const uint32_t st = stage0 + s * a.stage_bytes;
                const int32_t node = (int32_t)(n_begin + (int64_t)it * TN_BK);
                mbar_arrive_expect_tx(bar_full + 8 * s, tx);
                for (int32_t b = 0; b < nbA; b++)
                    tma_load_2d(st + b * BOX_BYTES, &tmA, bar_full + 8 * s, half * 128 + b * 32, node);
                for (int32_t b = 0; b < a.nbB; b++)
                    tma_load_2d(st + A_HI_BYTES + b * BOX_BYTES, &tmB, bar_full + 8 * s, (b_box0 + b) * 32, node);
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && (NCTA == 1 || half == 0)) {
            const uint32_t idesc = instr_desc(128 * NCTA, (uint32_t)a.N, 1, 1);
            StagePos pos;
            for (int32_t chunk = 0; chunk < nchunks; chunk++) {
                const uint32_t accb = chunk & 1, acc_ph = (chunk >> 1) & 1;
                if (NCTA == 2 && (a.debug & 8)) mbar_wait_cluster(bar_tempty + 8 * accb, acc_ph ^ 1);
                else mbar_wait(bar_tempty + 8 * accb, acc_ph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + accb * a.N;
                const int32_t nst = min(TN_CHUNK, iters - chunk * TN_CHUNK);
                for (int32_t jst = 0; jst < nst; jst++, pos.next(a.stages)) {
                    const uint32_t s = pos.s, ph = pos.ph;
                    if constexpr (NCTA == 2) {
                        // both CTAs' converters arrive here, each after its own stage landed
                        if (a.debug & 8) mbar_wait_cluster(bar_conv + 8 * s, ph);
                        else mbar_wait(bar_conv + 8 * s, ph);
                    } else {
                        mbar_wait(bar_full + 8 * s, ph);
                        mbar_wait(bar_conv + 8 * s, ph);
                    }
                    tc_fence_after();
                    const uint32_t st = stage0 + s * a.stage_bytes;
                    // MN-major, 32-byte-atom 128-byte swizzle: LBO = distance between 32-float column boxes,
                    // SBO = distance between groups of 4 node rows
                    const uint64_t a_hi = smem_desc(st, BOX_BYTES, 512, LAYOUT_SW128_BASE32B);
                    const uint64_t a_lo = smem_desc(st + a.hi_bytes, BOX_BYTES, 512, LAYOUT_SW128_BASE32B);
                    const uint64_t b_hi = smem_desc(st + A_HI_BYTES, BOX_BYTES, 512, LAYOUT_SW128_BASE32B);
                    const uint64_t b_lo = smem_desc(st + A_HI_BYTES + a.hi_bytes, BOX_BYTES, 512, LAYOUT_SW128_BASE32B);
#pragma unroll
                    for (uint32_t k = 0; k < TN_BK / 8; k++) {
                        const uint64_t adv = (uint64_t)(k * 1024 >> 4); // next group of 8 node rows
                        mma_issue<NCTA>(d_tmem, a_lo + adv, b_hi + adv, idesc, (jst | k) != 0);
                        mma_issue<NCTA>(d_tmem, a_hi + adv, b_lo + adv, idesc, 1);
                        mma_issue<NCTA>(d_tmem, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    mma_done<NCTA>(bar_empty + 8 * s);
                }
                mma_done<NCTA>(bar_tfull + 8 * accb);
            }
        }
    } else {
        const int t = threadIdx.x - 64; // 0..255
        const int q = warp & 3;         // TMEM lane quadrant this warp may read
        const int colhalf = (warp - 2) >> 2;
        const int32_t Nh = a.N >> 1;    // columns summed by this thread (multiple of 16, <= 128)
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; i++) acc[i] = 0.f;

        auto drain = [&](int32_t chunk) {
            const uint32_t accb = chunk & 1, acc_ph = (chunk >> 1) & 1;
            mbar_wait(bar_tfull + 8 * accb, acc_ph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + accb * a.N + colhalf * Nh;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (c * 16 < Nh) {
                    uint32_t r[16];
                    tmem_ld16(t_row + c * 16, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; e++) acc[c * 16 + e] += __uint_as_float(r[e]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (NCTA == 2) mbar_arrive_cluster_light(bar_tempty_l + 8 * accb);
                else mbar_arrive(bar_tempty + 8 * accb);
            }
        };

        int32_t drained = 0;
        const int32_t nboxes = nbA + a.nbB; // a box is 2 KB = 128 float4
        StagePos pos;
        for (int32_t it = 0; it < iters; it++, pos.next(a.stages)) {
            const uint32_t s = pos.s, ph = pos.ph;
            mbar_wait(bar_full + 8 * s, ph);
            uint8_t *st = gbase + CTRL_BYTES + (size_t)s * a.stage_bytes + (t & 127) * 16;
            // thread t converts vector (t & 127) of boxes (t >> 7), (t >> 7) + 2, ...: up to 6 boxes (4 of A + 8 of B over
            // 256 threads), unrolled so that the shared-memory loads of all of them are in flight together (ncu: the
            // rolled loop made this warp role a ~100-cycle dependent chain per box, 300 instructions per warp and stage)
#pragma unroll
            for (int j0 = 0; j0 < 6; j0 += 2) { // two boxes at a time (more would spill: 168 registers is the cap of 3 warps per scheduler): loads first, then the conversions and stores
                float4 raw[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int32_t b = (t >> 7) + 2 * (j0 + j);
                    const uint32_t off = b < nbA ? b * BOX_BYTES : A_HI_BYTES + (b - nbA) * BOX_BYTES;
                    if (b < nboxes) raw[j] = *reinterpret_cast<const float4 *>(st + off);
                }
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    const int32_t b = (t >> 7) + 2 * (j0 + j);
                    const uint32_t off = b < nbA ? b * BOX_BYTES : A_HI_BYTES + (b - nbA) * BOX_BYTES;
                    if (b < nboxes) {
                        float4 *hp = reinterpret_cast<float4 *>(st + off);
                        if constexpr (TRUNC) split4_trunc(raw[j], hp + a.hi_bytes / 16);
                        else split4_from(raw[j], hp, hp + a.hi_bytes / 16);
                    }
                }
            }
            if constexpr (NCTA == 2) {
                pair_publish(bar_conv_l + 8 * s, lane, a.debug & 8);
            } else {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_conv + 8 * s);
            }
            // chunk `drained` is complete once the pipeline has moved a.stages stages past its end
            if (it + 1 == (drained + 1) * TN_CHUNK + a.stages) drain(drained++);
        }
        while (drained < nchunks) drain(drained++);

        float *orow = a.partial + (size_t)split * a.part_stride + (size_t)(half * 128 + q * 32 + lane) * a.N +
                      colhalf * Nh;
#pragma unroll
        for (int c = 0; c < 32; c++)
            if (c * 4 < Nh)
                *reinterpret_cast<float4 *>(orow + c * 4) =
                    make_float4(acc[c * 4], acc[c * 4 + 1], acc[c * 4 + 2], acc[c * 4 + 3]);
    }

    tc_fence_before();
    __syncwarp();
    if constexpr (NCTA == 2) cluster_sync_all();
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, a.tmem_cols);
    }
}

// C[r, c] = sum over CTAs (ascending) of partial[cta][r][c]
__global__ void tn_reduce_kernel(const float *__restrict__ partial, int32_t n_parts, int64_t part_stride, int32_t N,
                                 int32_t K1, int32_t K2, float *__restrict__ C, int64_t ldc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)K1 * K2) return;
    const int32_t r = (int32_t)(i / K2), c = (int32_t)(i % K2);
    const float *p = partial + (size_t)r * N + c;
    float s = 0.f;
    for (int32_t z = 0; z < n_parts; z++) s += p[(size_t)z * part_stride];
    C[(int64_t)r * ldc + c] = s;
}

// out[c] = sum over the epilogue warps' partials (ascending) of part[p][c], c < F
__global__ void rows_colsum_final_kernel(const float *__restrict__ part, int32_t n_parts, int32_t F, float *__restrict__ out) {
    const int32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= F) return;
    float s = 0.f;
    for (int32_t p = 0; p < n_parts; p++) s += part[(size_t)p * 256 + c];
    out[c] = s;
}

// weights -> [2*Npad, Kpad]: rows [0,Npad) = hi, rows [Npad,2Npad) = lo of Bt[n][k], zero padded.
// transpose = 0: Bt[n][k] = B[n*ldb + k] (NT);  1: Bt[n][k] = B[k*ldb + n] (NN)
__global__ void prep_weights_kernel(const float *__restrict__ B, int64_t ldb, int32_t N, int32_t K, int transpose,
                                    float *__restrict__ out, int32_t Npad, int32_t Kpad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)Npad * Kpad) return;
    const int32_t n = (int32_t)(i / Kpad), k = (int32_t)(i % Kpad);
    float v = 0.f;
    if (n < N && k < K) v = transpose ? B[(int64_t)k * ldb + n] : B[(int64_t)n * ldb + k];
    float hi = __uint_as_float(rna_tf32(v));
    if (!isfinite(hi)) hi = v; // keep inf/nan (and values that would round up to inf) in the hi part alone
    const float lo = isfinite(hi) ? __uint_as_float(rna_tf32(v - hi)) : 0.f;
    out[i] = hi;
    out[(int64_t)Npad * Kpad + i] = lo;
}


// ------------------------------------------------------------------------------------------------ TF32 peak probe
// Dense TF32 tensor-core peak of this GPU, measured (BASELINE.md §3 asks for it as the compute-side roofline
// denominator of the dense transforms): every CTA issues a long dependent-free stream of 128 x 256 x 8
// tcgen05.mma kind::tf32 on one zero-filled shared-memory operand pair (the operands are re-read from shared memory by
// every instruction, as in the real kernels), no global traffic.
__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(int iters) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw_u32 = smem_u32(smem_raw);
    const uint32_t base = (raw_u32 + 1023u) & ~1023u;
    uint8_t *gbase = smem_raw + (base - raw_u32);
    const uint32_t bar = base, tmem_slot = base + 16, a_tile = base + 1024, b_tile = a_tile + 128 * 128;
    for (uint32_t i = threadIdx.x; i < (128 * 128 + 256 * 128) / 16; i += blockDim.x)
        reinterpret_cast<float4 *>(gbase + 1024)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t *>(gbase + 16);
    if (warp == 0 && lane == 0) {
        const uint32_t idesc = instr_desc(128, 256, 0, 0);
        const uint64_t da = smem_desc(a_tile, 16, 1024), db = smem_desc(b_tile, 16, 1024);
        uint32_t ph = 0;
        for (int it = 0; it < iters; it++) {
            for (uint32_t k = 0; k < 4; k++) mma_tf32(tmem_base, da + (uint64_t)(k * 32 >> 4), db + (uint64_t)(k * 32 >> 4), idesc, 1);
            if ((it & 63) == 63 || it + 1 == iters) { // bound the number of MMAs in flight
                mma_commit(bar);
                mbar_wait(bar, ph);
                ph ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 256); }
}

// ------------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// 2-D FP32 row-major view [rows, cols] with leading dimension ld; box = {32 floats, box_rows}, 128-byte swizzle,
// out-of-bounds elements read as zero
static int make_map(CUtensorMap *tm, const float *base, int64_t rows, int64_t cols, int64_t ld, uint32_t box_rows,
                    CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B, uint32_t box_cols = 32) {
    EncodeTiledFn fn = encode_fn();
    GNN_REQUIRE(fn, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)base, gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GNN_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                (long long)rows, (long long)cols, (long long)ld);
    return 0;
}

static bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }
static uint32_t pow2_cols(uint32_t c) {
    uint32_t p = 32;
    while (p < c) p <<= 1;
    return p;
}
constexpr uint32_t SMEM_MAX = 227 * 1024;

// GNN_GEMM_SPLIT: 1 (default) = truncating hi/lo split of the streamed operand, 0 = round-to-nearest (two cvt.rna per
// element).  Measured (profiles/r2b_gemm_pair.md): the eight GEMMs of a products-shaped step 11.00 -> 9.54 ms, the TN
// kernels (both operands streamed and split) 0.97 / 2.02 / 1.84 -> 0.70 / 1.44 / 1.41 ms; worst error against the fp64
// oracle 2.6e-6 -> 2.8e-6 (isolated products), 4.5e-6 -> 5.3e-6 (dW_1 of the full train step); bar 1e-5.
static int split_mode() {
    const char *e = getenv("GNN_GEMM_SPLIT");
    return e ? atoi(e) : 1;
}

// GNN_GEMM_PAIR: bit 0 = CTA pairs in the rows kernel (NT, NN), bit 1 = in the TN kernel; default both
static int pair_mask() {
    const char *e = getenv("GNN_GEMM_PAIR");
    return e ? atoi(e) : 3;
}

// CTA pairs of `fn` (1 CTA per SM, cluster of 2) that can be resident at once: the persistent grids are sized to it.
// Falls back to SMs / 2 when the occupancy query is not answered.
static int pair_capacity(gnn_ctx *ctx, const void *fn, cudaLaunchConfig_t *cfg) {
    struct Entry { const void *fn; int device; size_t smem; int n; };
    static Entry cache[16];
    static int cached = 0;
    for (int i = 0; i < cached; i++)
        if (cache[i].fn == fn && cache[i].device == ctx->device && cache[i].smem == cfg->dynamicSmemBytes) return cache[i].n;
    cfg->gridDim = dim3(2 * (unsigned)(ctx->sm_count / 2));
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fn, cfg) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = ctx->sm_count / 2;
    }
    if (n > ctx->sm_count / 2) n = ctx->sm_count / 2;
    if (n < 1) n = 1;
    if (getenv("GNN_GEMM_DEBUG_PRINT")) fprintf(stderr, "gemm_tc: %d CTA pairs resident (smem %zu)\n", n, cfg->dynamicSmemBytes);
    if (cached < 16) cache[cached++] = Entry{fn, ctx->device, cfg->dynamicSmemBytes, n};
    return n;
}

// C[M, N] (N <= 256) = A[M,K] * Bt^T with Bt given through `transpose` as in prep_weights_kernel
static int rows_gemm(gnn_ctx *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B,
                     int64_t ldb, int transpose, float *C, int64_t ldc, const float *bias, int relu, const float *mask,
                     int64_t ldm, float *colsum_out = nullptr) {
    const int32_t Npad = (int32_t)round_up(N, 16), Kpad = (int32_t)round_up(K, BK);
    const bool lsu_epi = (N % 4 != 0) || (getenv("GNN_GEMM_EPI") && !strcmp(getenv("GNN_GEMM_EPI"), "lsu"));
    if (colsum_out && lsu_epi) return -1; // the fused column sums live in the TMA epilogue
    void *ws = nullptr;
    const size_t bs_bytes = (size_t)round_up((int64_t)2 * Npad * Kpad * 4, 256);
    const size_t part_bytes = colsum_out ? (size_t)ctx->sm_count * 4 * 256 * 4 : 0;
    GNN_TRY(ctx->workspace(bs_bytes + part_bytes, &ws));
    float *Bs = (float *)ws;
    {
        const int64_t n = (int64_t)Npad * Kpad;
        prep_weights_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(B, ldb, N, K, transpose, Bs, Npad, Kpad);
        GNN_LAUNCHED(ctx);
    }
    // A stage holds the A tile twice (hi, lo) and both halves of the weight k-block.  With 32-float k-blocks a
    // 256-wide output leaves room for only 2 stages (96 KB each) and the TMA -> convert -> MMA chain of a stage
    // cannot overlap enough; 16-float k-blocks (64-byte swizzle) halve the stage and double the depth.
    RowsArgs a;
    // CTA pairs (cta_group::2, 256-row tiles) unless GNN_GEMM_PAIR clears bit 0: each CTA stages half of the weights' output columns
    // — which pays where the weight traffic is large: wide outputs whose reduction is too long to keep the weights
    // resident (measured, ms per launch, single -> pair: K=256,N=256 1.70 -> 1.38; K=100,N=256 0.93 -> 0.93;
    // K=256,N=47 1.08 -> 1.06; K=47,N=256 (resident) 1.31 -> 1.41)
    const int ncta = !(pair_mask() & 1) || ctx->sm_count < 2 || M <= TILE_M || Npad < 128 || Kpad <= 64 ? 1 : 2;
    const uint32_t nb = (uint32_t)Npad / ncta; // weight rows (output columns) staged per CTA: a multiple of 8
    // TMA stores clip the tensor edge in 16-byte units (observed: column 47 of a 47-wide, ld 48 output was written),
    // so an output whose width is not a multiple of 4 keeps the per-lane store epilogue with its exact guards
    a.tma_epi = lsu_epi ? 0 : 1;
    a.colsum_part = colsum_out ? (float *)((uint8_t *)ws + bs_bytes) : nullptr;
    a.epi_bytes = a.tma_epi ? (mask ? 65536u : 32768u) : STAGING_BYTES;
    const uint32_t fixed = 1024 /*align slack*/ + CTRL_BYTES + a.epi_bytes;
    a.bk = BK;
    // Short reductions: the whole split weight matrix (2 * Npad * Kpad floats) stays resident in shared memory and
    // only A is streamed (otherwise every k-block of every tile re-fetches 2 * Npad * bk weights from L2).
    const uint32_t b_all = 2u * nb * (uint32_t)Kpad * 4u;
    const bool resident = Kpad <= 64 && b_all + 3 * (2 * A_TILE_BYTES) <= SMEM_MAX - fixed && !getenv("GNN_GEMM_NO_RESIDENT");
    // 32-float k-blocks unless that leaves fewer than 3 pipeline stages (measured: N=47, K=256 runs 1.06 ms with
    // four 32-float stages and 1.50 ms with 16-float ones; a 256-wide output only fits 16-float stages)
    const uint32_t stage32 = 2 * A_TILE_BYTES + (resident ? 0 : 2 * nb * 128);
    if ((SMEM_MAX - fixed - (resident ? b_all : 0)) / stage32 < 3) a.bk = 16;
    if (getenv("GNN_GEMM_BK")) a.bk = atoi(getenv("GNN_GEMM_BK")) == 16 ? 16 : 32;
    a.a_bytes = (uint32_t)TILE_M * a.bk * 4;
    a.b_bytes = nb * a.bk * 4;
    a.b_resident = resident ? b_all : 0;
    a.debug = getenv("GNN_GEMM_DEBUG") ? (uint32_t)atoi(getenv("GNN_GEMM_DEBUG")) : 0;
    const bool trunc = split_mode() != 0;
    a.stage_bytes = 2 * a.a_bytes + (resident ? 0 : 2 * a.b_bytes);
    CUtensorMap tmA, tmB;
    const CUtensorMapSwizzle sw = a.bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    GNN_TRY(make_map(&tmA, A, M, K, lda, TILE_M, sw, (uint32_t)a.bk));
    GNN_TRY(make_map(&tmB, Bs, 2 * (int64_t)Npad, Kpad, Kpad, nb, sw, (uint32_t)a.bk));

    a.M = M; a.N = N; a.Npad = Npad; a.kblocks = Kpad / a.bk;
    a.num_tiles = (int32_t)ceil_div(M, TILE_M * ncta);
    int stages = (int)((SMEM_MAX - fixed - a.b_resident) / a.stage_bytes);
    if (stages > 8) stages = 8;
    GNN_REQUIRE(stages >= 2, "gemm_tc: tile does not fit shared memory");
    a.stages = stages;
    a.acc_stride = pow2_cols((uint32_t)Npad);
    a.tmem_cols = 2 * a.acc_stride;
    a.C = C; a.ldc = ldc; a.bias = bias; a.relu = relu; a.mask = mask; a.ldm = ldm;
    const uint32_t smem = fixed + a.b_resident + (uint32_t)stages * a.stage_bytes;
    static uint64_t attr_set = 0; // function attributes are per device: one bit per device ordinal
    if (!(attr_set >> (ctx->device & 63) & 1)) {
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_rows_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_rows_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_rows_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_rows_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        attr_set |= 1ull << (ctx->device & 63);
    }
    CUtensorMap tmC, tmM;
    GNN_TRY(make_map(&tmC, C, M, N, ldc, 32));
    if (mask) GNN_TRY(make_map(&tmM, mask, M, N, ldm, 32));
    else tmM = tmC;
    int grid = a.num_tiles < ctx->sm_count ? a.num_tiles : ctx->sm_count; // persistent: one CTA per SM
    if (ncta == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(ROWS_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int pairs = pair_capacity(ctx, (const void *)tc_rows_kernel<2, false>, &cfg); // one CTA pair per TPC
        if (pairs > a.num_tiles) pairs = a.num_tiles;
        grid = 2 * pairs;
        cfg.gridDim = dim3((unsigned)grid);
        if (trunc) GNN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_rows_kernel<2, true>, tmA, tmB, tmC, tmM, a));
        else GNN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_rows_kernel<2, false>, tmA, tmB, tmC, tmM, a));
    } else if (trunc) {
        tc_rows_kernel<1, true><<<grid, ROWS_THREADS, smem, ctx->stream>>>(tmA, tmB, tmC, tmM, a);
    } else {
        tc_rows_kernel<1, false><<<grid, ROWS_THREADS, smem, ctx->stream>>>(tmA, tmB, tmC, tmM, a);
    }
    GNN_LAUNCHED(ctx);
    if (colsum_out) {
        rows_colsum_final_kernel<<<(unsigned)ceil_div(N, 128), 128, 0, ctx->stream>>>(a.colsum_part, grid * 4, N, colsum_out);
        GNN_LAUNCHED(ctx);
    }
    return 0;
}

static uint64_t tn_attr_set = 0; // per-device function attributes of the TN kernels: one bit per device ordinal
// C[K1, K2 (<=256)] = A[M,K1]^T B[M,K2]
static int tn_gemm(gnn_ctx *ctx, int64_t M, int32_t K1, int32_t K2, const float *A, int64_t lda, const float *B,
                   int64_t ldb, float *C, int64_t ldc) {
    if (!(tn_attr_set >> (ctx->device & 63) & 1)) {
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_tn_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_tn_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_tn_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        GNN_CHECK_CUDA(cudaFuncSetAttribute(tc_tn_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
        tn_attr_set |= 1ull << (ctx->device & 63);
    }
    TnArgs a;
    a.M = M;
    a.K1 = K1;
    const int32_t halves = (int32_t)ceil_div(K1, 128);
    // two halves = one CTA pair (cta_group::2) unless GNN_GEMM_PAIR clears bit 1: each CTA then stages half of B's columns
    const int ncta = halves == 2 && (pair_mask() & 2) && ctx->sm_count >= 2 ? 2 : 1;
    a.N = (int32_t)round_up(K2, 32 * ncta);
    a.nbB = a.N / (32 * ncta);
    a.hi_bytes = (uint32_t)(4 + a.nbB) * BOX_BYTES;
    a.stage_bytes = 2 * a.hi_bytes;
    const uint32_t fixed = 1024 + CTRL_BYTES;
    int stages = (int)((SMEM_MAX - fixed) / a.stage_bytes);
    if (stages > (ncta == 2 ? 8 : 6)) stages = ncta == 2 ? 8 : 6;
    GNN_REQUIRE(stages >= 2, "gemm_tc: tile does not fit shared memory");
    a.stages = stages;
    a.tmem_cols = pow2_cols((uint32_t)(2 * a.N));
    int64_t splits = ctx->sm_count / halves;
    if (ncta == 2) { // one CTA pair per TPC that can hold one
        cudaLaunchConfig_t q = {};
        q.blockDim = dim3(TN_THREADS);
        q.dynamicSmemBytes = fixed + (uint32_t)stages * a.stage_bytes;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa;
        q.numAttrs = 1;
        splits = pair_capacity(ctx, (const void *)tc_tn_kernel<2, false>, &q);
    }
    if (splits < 1) splits = 1;
    const int64_t max_splits = ceil_div(M, TN_BK);
    if (splits > max_splits) splits = max_splits;
    a.nodes_per_cta = round_up(ceil_div(M, splits), TN_BK);
    splits = ceil_div(M, a.nodes_per_cta);
    a.part_stride = (int64_t)halves * 128 * a.N;
    void *ws = nullptr;
    GNN_TRY(ctx->workspace((size_t)splits * a.part_stride * 4, &ws));
    a.partial = (float *)ws;
    a.debug = getenv("GNN_GEMM_DEBUG") ? (uint32_t)atoi(getenv("GNN_GEMM_DEBUG")) : 0;
    const bool trunc = split_mode() != 0;
    CUtensorMap tmA, tmB;
    GNN_TRY(make_map(&tmA, A, M, K1, lda, TN_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    GNN_TRY(make_map(&tmB, B, M, K2, ldb, TN_BK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B));
    const uint32_t smem = fixed + (uint32_t)stages * a.stage_bytes;
    if (ncta == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (unsigned)splits);
        cfg.blockDim = dim3(TN_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (trunc) GNN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_tn_kernel<2, true>, tmA, tmB, a));
        else GNN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tc_tn_kernel<2, false>, tmA, tmB, a));
    } else if (trunc) {
        tc_tn_kernel<1, true><<<dim3((unsigned)splits, (unsigned)halves), TN_THREADS, smem, ctx->stream>>>(tmA, tmB, a);
    } else {
        tc_tn_kernel<1, false><<<dim3((unsigned)splits, (unsigned)halves), TN_THREADS, smem, ctx->stream>>>(tmA, tmB, a);
    }
    GNN_LAUNCHED(ctx);
    const int64_t n = (int64_t)K1 * K2;
    tn_reduce_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(a.partial, (int32_t)splits, a.part_stride, a.N,
                                                                        K1, K2, C, ldc);
    GNN_LAUNCHED(ctx);
    return 0;
}

} // namespace tc

} // namespace gnn
extern "C" int gnn_tf32_peak_probe(gnn_ctx_t *ctx, double *tflops_h) {
    using namespace gnn;
    GNN_REQUIRE(ctx && tflops_h, "gnn_tf32_peak_probe: NULL argument");
    const int iters = 20000, grid = ctx->sm_count;
    const size_t smem = 1024 + 1024 + 128 * 128 + 256 * 128;
    GNN_CHECK_CUDA(cudaFuncSetAttribute(tc::tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t a, b;
    GNN_CHECK_CUDA(cudaEventCreate(&a));
    GNN_CHECK_CUDA(cudaEventCreate(&b));
    tc::tf32_peak_kernel<<<grid, 128, smem, ctx->stream>>>(iters / 10); // warm-up
    GNN_CHECK_CUDA(cudaEventRecord(a, ctx->stream));
    tc::tf32_peak_kernel<<<grid, 128, smem, ctx->stream>>>(iters);
    GNN_LAUNCHED(ctx);
    GNN_CHECK_CUDA(cudaEventRecord(b, ctx->stream));
    GNN_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    GNN_CHECK_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    *tflops_h = (double)grid * iters * 4.0 * 2.0 * 128 * 256 * 8 / (ms * 1e-3) / 1e12;
    return 0;
}
namespace gnn {

// The TMA path needs 16-byte aligned rows; anything else goes back to the FP32 FMA kernel (return -1).
// The rows kernel keeps a whole reduction in one TMEM accumulator, whose truncating adds cost ~2e-6 of relative
// error per 256 of K (measured): reductions longer than 512 stay on the FP32 FMA kernel to hold the 1e-5 contract.
constexpr int32_t TC_MAX_K = 512;
int gemm_tc_nt(gnn_ctx *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B, int64_t ldb,
               float *C, int64_t ldc, const float *bias, int relu) {
    if (!tc::aligned16(A) || !tc::aligned16(C) || (lda & 3) || (ldc & 3)) return -1;
    if (M >= (1ll << 31) - 512 || K > TC_MAX_K) return -1;
    for (int32_t n0 = 0; n0 < N; n0 += 256) { // wider outputs: column panels of 256
        const int32_t nn = N - n0 < 256 ? N - n0 : 256;
        GNN_TRY(tc::rows_gemm(ctx, M, nn, K, A, lda, B + (int64_t)n0 * ldb, ldb, 0, C + n0, ldc, bias ? bias + n0 : nullptr,
                              relu, nullptr, 0));
    }
    return 0;
}

// colsum_out != NULL: also the column sums of C (the bias gradient of the layer below), from the same epilogue; only for
// a single 256-column panel with the TMA epilogue — otherwise -1 before anything is launched
int gemm_tc_nn(gnn_ctx *ctx, int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B, int64_t ldb,
               float *C, int64_t ldc, const float *mask, int64_t ldm, float *colsum_out) {
    if (!tc::aligned16(A) || !tc::aligned16(C) || (lda & 3) || (ldc & 3)) return -1;
    if (colsum_out && (N > 256 || (N & 3))) return -1;
    if (mask && (!tc::aligned16(mask) || (ldm & 3))) return -1;
    if (M >= (1ll << 31) - 512 || K > TC_MAX_K) return -1;
    for (int32_t n0 = 0; n0 < N; n0 += 256) {
        const int32_t nn = N - n0 < 256 ? N - n0 : 256;
        GNN_TRY(tc::rows_gemm(ctx, M, nn, K, A, lda, B + n0, ldb, 1, C + n0, ldc, nullptr, 0, mask ? mask + n0 : nullptr,
                              ldm, colsum_out));
    }
    return 0;
}

int gemm_tc_tn(gnn_ctx *ctx, int64_t M, int32_t K1, int32_t K2, const float *A, int64_t lda, const float *B,
               int64_t ldb, float *C, int64_t ldc) {
    if (!tc::aligned16(A) || !tc::aligned16(B) || (lda & 3) || (ldb & 3)) return -1;
    if (M >= (1ll << 31)) return -1;
    for (int32_t c0 = 0; c0 < K2; c0 += 256) { // wider outputs: column panels of 256
        const int32_t kc = K2 - c0 < 256 ? K2 - c0 : 256;
        GNN_TRY(tc::tn_gemm(ctx, M, K1, kc, A, lda, B + c0, ldb, C + c0, ldc));
    }
    return 0;
}

} // namespace gnn
