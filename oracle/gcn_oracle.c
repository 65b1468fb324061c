/* gcn_oracle.c — CPU restatement of walexi/gnn.cpp's GCN hot path (plain C, ctypes-callable).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this.  The product path (gnn.cpp_b200/, include/) never does.
 *
 * PARITY PINNING: every function below is checked against outputs of the reference itself
 * (oracle/_ref/ref_gcn, built by oracle/build_ref.sh from /root/reference) through the committed
 * fixtures in tests/golden/ (tests/test_oracle_vs_reference.py).  Exceptions, where the reference has
 * no working implementation ("parity unpinned" in the reference, pinned by this file only):
 *   - orc_csc_from_csr, orc_partition_*  (no sparse formats / no multi-GPU in the reference)
 *   - orc_sgd_step                       (nn::SGD::step segfaults, SURVEY.md bug B4; torch semantics
 *                                         per include/nn.h:165-167)
 *   - dZ of the loss                     (nn::cross_entropy_loss backward throws, bug B3; analytic
 *                                         (softmax - onehot)/N from nn::softmax, src/nn.cpp:270-278)
 *
 * "Reference order" = the accumulation order the reference arithmetic uses, so that this file and
 * ref_gcn agree to the last bit where possible:
 *   - matmul dot products: k DESCENDING, sequential fp32, no FMA   include/functional.h:432-439 +
 *     libstdc++ _Expr::sum (valarray_after.h)
 *   - sum along a dim / whole tensor: ASCENDING sequential fp32     include/functional.h:266-296
 * Compile with -ffp-contract=off.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------------
 * Synthetic inputs.  Counter-based splitmix64 so numpy (gnn.cpp_b200/synth.py), this
 * file and the C++ driver generate bit-identical problems.
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static inline uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t i) {
    return mix64(mix64(seed * 0x9E3779B97F4A7C15ULL + stream * 0xD1B54A32D192ED03ULL) + i);
}
ORC_API uint64_t orc_hash3(uint64_t seed, uint64_t stream, uint64_t i) { return hash3(seed, stream, i); }

/* U[lo,hi) floats: 24 random bits -> exact float in [0,1) -> lo + (hi-lo)*u in fp32 */
ORC_API void orc_synth_uniform(uint64_t seed, uint64_t stream, int64_t n, float lo, float hi, float *out) {
    const float span = hi - lo;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) {
        float u = (float)(hash3(seed, stream, (uint64_t)i) >> 40) * 0x1p-24f;
        float t = span * u;
        out[i] = lo + t;
    }
}
ORC_API void orc_synth_labels(uint64_t seed, uint64_t stream, int64_t n, int32_t C, int32_t *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; i++) out[i] = (int32_t)(hash3(seed, stream, (uint64_t)i) % (uint64_t)C);
}
static inline int32_t synth_endpoint(uint64_t h, int32_t N, int powerlaw, uint64_t pa, uint64_t pb) {
    if (!powerlaw) return (int32_t)(h % (uint64_t)N);
    /* Chung-Lu style skew: id = floor(N * u^1.5), then an affine bijection scatters hot ids */
    double x = (double)(h >> 32) * 0x1p-32;
    double s = sqrt(x);
    double t = x * s;
    int64_t id = (int64_t)(t * (double)N);
    if (id >= N) id = N - 1;
    return (int32_t)((pa * (uint64_t)id + pb) % (uint64_t)N);
}
static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t t = a % b; a = b; b = t; } return a; }
/* E directed entries = npairs undirected pairs, symmetrised: src=[u,v], dst=[v,u].  Self loops and
 * duplicates are left in on purpose (the structure build must collapse them like graph.cpp:21-75). */
ORC_API void orc_synth_edges(uint64_t seed, int64_t E, int32_t N, int powerlaw, int32_t *src, int32_t *dst) {
    int64_t npairs = E / 2;
    uint64_t pa = 0x9E3779B1ULL % (uint64_t)N;
    if (pa == 0) pa = 1;
    while (gcd_u64(pa, (uint64_t)N) != 1) pa++;
    uint64_t pb = 0x7F4A7C15ULL % (uint64_t)N;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < npairs; k++) {
        int32_t u = synth_endpoint(hash3(seed, 1, (uint64_t)k), N, powerlaw, pa, pb);
        int32_t v = synth_endpoint(hash3(seed, 2, (uint64_t)k), N, powerlaw, pa, pb);
        src[k] = u; dst[k] = v;
        src[npairs + k] = v; dst[npairs + k] = u;
    }
    if (E & 1) { /* odd E: one extra directed entry */
        src[E - 1] = synth_endpoint(hash3(seed, 1, (uint64_t)npairs), N, powerlaw, pa, pb);
        dst[E - 1] = synth_endpoint(hash3(seed, 2, (uint64_t)npairs), N, powerlaw, pa, pb);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Structure: COO -> row-major sorted, de-duplicated CSR with the diagonal forced.
 *   edge_to_adj_mat  (src/graph.cpp:21-44):  A[src*N+dst] = 1  — assignment, duplicates collapse,
 *                                            row = edge_index[0], col = edge_index[1]
 *   fill_diagonal_   (include/tensor.h:806-817): diagonal := fill (0 removes loops, 1 adds them)
 *   adj_to_edge_list (src/graph.cpp:46-67):  row-major scan, keeps int(a)!=0  => rows ascending,
 *                                            cols ascending inside a row
 * fill_mode: 0 = diagonal removed, 1 = diagonal present on every row, 2 = leave as given.
 * Returns nnz; rowptr has N+1 int64 entries; colidx capacity must be >= E + N.
 * ---------------------------------------------------------------------------------------------- */
static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return (x > y) - (x < y);
}
static void radix_sort_u64(uint64_t *keys, int64_t n, int bits) {
    if (n < 4096) { qsort(keys, (size_t)n, 8, cmp_u64); return; }
    uint64_t *tmp = (uint64_t *)malloc((size_t)n * 8);
    uint64_t *a = keys, *b = tmp;
    for (int shift = 0; shift < bits; shift += 16) {
        int64_t *cnt = (int64_t *)calloc(65537, sizeof(int64_t));
        for (int64_t i = 0; i < n; i++) cnt[((a[i] >> shift) & 0xFFFF) + 1]++;
        for (int i = 0; i < 65536; i++) cnt[i + 1] += cnt[i];
        for (int64_t i = 0; i < n; i++) b[cnt[(a[i] >> shift) & 0xFFFF]++] = a[i];
        free(cnt);
        uint64_t *t = a; a = b; b = t;
    }
    if (a != keys) memcpy(keys, a, (size_t)n * 8);
    free(tmp);
}
ORC_API int64_t orc_csr_build(const int32_t *src, const int32_t *dst, int64_t E, int32_t N, int fill_mode,
                              int64_t *rowptr, int32_t *colidx) {
    int64_t cap = E + (fill_mode == 1 ? N : 0);
    uint64_t *keys = (uint64_t *)malloc((size_t)(cap > 0 ? cap : 1) * 8);
    int64_t m = 0;
    for (int64_t e = 0; e < E; e++) {
        if (fill_mode == 0 && src[e] == dst[e]) continue;
        keys[m++] = ((uint64_t)(uint32_t)src[e] << 32) | (uint32_t)dst[e];
    }
    if (fill_mode == 1)
        for (int32_t i = 0; i < N; i++) keys[m++] = ((uint64_t)(uint32_t)i << 32) | (uint32_t)i;
    radix_sort_u64(keys, m, 64);
    int64_t nnz = 0;
    memset(rowptr, 0, (size_t)(N + 1) * 8);
    for (int64_t i = 0; i < m; i++) {
        if (i > 0 && keys[i] == keys[i - 1]) continue;
        int32_t r = (int32_t)(keys[i] >> 32);
        colidx[nnz++] = (int32_t)(keys[i] & 0xFFFFFFFFu);
        rowptr[r + 1]++;
    }
    for (int32_t r = 0; r < N; r++) rowptr[r + 1] += rowptr[r];
    free(keys);
    return nnz;
}

/* Weighted adjacency (graph::edge_to_adj_mat with edge_attr, src/graph.cpp:21-44): A[src][dst] = w by ASSIGNMENT in
 * edge order, so for duplicate (src, dst) pairs the LAST weight wins; fill_mode 1 then forces the diagonal to 1
 * (tensor::fill_diagonal_(1), include/tensor.h:806-817: the A + I of mode B), fill_mode 2 leaves it as given.
 * Output: CSR (rows ascending, columns ascending) with the raw weights val0.  Entries whose final weight is exactly 0
 * stay as explicit zeros (they are zeros of the dense matrix and contribute nothing). */
typedef struct { uint64_t key; int64_t idx; } orc_item;
static int orc_item_cmp(const void *a, const void *b) {
    const orc_item *x = (const orc_item *)a, *y = (const orc_item *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
ORC_API int64_t orc_csr_build_weighted(const int32_t *src, const int32_t *dst, const float *w, int64_t E, int32_t N,
                                       int fill_mode, int64_t *rowptr, int32_t *colidx, float *val0) {
    const int64_t m = E + (fill_mode == 1 ? N : 0);
    orc_item *it = (orc_item *)malloc((size_t)(m > 0 ? m : 1) * sizeof(orc_item));
    for (int64_t e = 0; e < E; e++) { it[e].key = ((uint64_t)(uint32_t)src[e] << 32) | (uint32_t)dst[e]; it[e].idx = e; }
    if (fill_mode == 1)
        for (int32_t i = 0; i < N; i++) { it[E + i].key = ((uint64_t)(uint32_t)i << 32) | (uint32_t)i; it[E + i].idx = E + i; }
    qsort(it, (size_t)m, sizeof(orc_item), orc_item_cmp);
    int64_t nnz = 0;
    memset(rowptr, 0, (size_t)(N + 1) * 8);
    for (int64_t i = 0; i < m; i++) {
        if (i + 1 < m && it[i + 1].key == it[i].key) continue; /* not the last write to this position */
        colidx[nnz] = (int32_t)(it[i].key & 0xFFFFFFFFu);
        val0[nnz] = it[i].idx < E ? w[it[i].idx] : 1.0f;
        rowptr[(int32_t)(it[i].key >> 32) + 1]++;
        nnz++;
    }
    for (int32_t r = 0; r < N; r++) rowptr[r + 1] += rowptr[r];
    free(it);
    return nnz;
}
/* Weighted mode B: deg = A->sum(-1,true) (ascending fp32 row sum, src/graph.cpp:178 / functional.h:266-296),
 * dinv = pow(deg,-0.5), val = (A*dinv)*dinv^T (functional.h:189-213). */
ORC_API void orc_degree_norm_weighted(int32_t N, const int64_t *rowptr, const int32_t *colidx, const float *val0,
                                      float *degf, float *dinv, float *val) {
    for (int32_t r = 0; r < N; r++) {
        float s = 0.0f;
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) s = s + val0[k];
        degf[r] = s;
        dinv[r] = powf(s, -0.5f);
    }
    for (int32_t r = 0; r < N; r++)
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) val[k] = (val0[k] * dinv[r]) * dinv[colidx[k]];
}

/* CSR -> CSC (= CSR of the transpose), rows ascending inside each column, and the permutation
 * perm[k] = CSR position of CSC entry k.  No counterpart in the reference (it transposes the dense
 * matrix, include/functional.h:330-357; include/operation.h:526-528). */
ORC_API void orc_csc_from_csr(int32_t N, const int64_t *rowptr, const int32_t *colidx, int64_t *colptr,
                              int32_t *rowidx, int64_t *perm) {
    int64_t nnz = rowptr[N];
    memset(colptr, 0, (size_t)(N + 1) * 8);
    for (int64_t k = 0; k < nnz; k++) colptr[colidx[k] + 1]++;
    for (int32_t c = 0; c < N; c++) colptr[c + 1] += colptr[c];
    int64_t *cur = (int64_t *)malloc((size_t)N * 8);
    memcpy(cur, colptr, (size_t)N * 8);
    for (int32_t r = 0; r < N; r++)
        for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) {
            int64_t p = cur[colidx[k]]++;
            rowidx[p] = r;
            perm[p] = k;
        }
    free(cur);
}

/* Degree / D^-1/2 / edge values.
 *   deg  = rowsum(A0 + I)                 src/graph.cpp:178 (sum(-1,true) + 1), exact integer in fp32
 *   dinv = std::pow(deg, -0.5f)           src/graph.cpp:183 -> include/functional.h:253
 *   val  = (1 * dinv[r]) * dinv[c]        mode-B composition (A*dinv)*dinv^T, functional.h:189-213 */
ORC_API void orc_degree_norm(int32_t N, const int64_t *rowptr, const int32_t *colidx, int32_t *deg, float *dinv,
                             float *val) {
    for (int32_t r = 0; r < N; r++) {
        deg[r] = (int32_t)(rowptr[r + 1] - rowptr[r]);
        dinv[r] = powf((float)deg[r], -0.5f);
    }
    if (val)
        for (int32_t r = 0; r < N; r++)
            for (int64_t k = rowptr[r]; k < rowptr[r + 1]; k++) val[k] = (1.0f * dinv[r]) * dinv[colidx[k]];
}

/* ------------------------------------------------------------------------------------------------
 * Aggregation  Y[N,F] = Ahat * P   (src/graph.cpp:208 -> include/functional.h:398-441).
 * Reference order: for each output, columns DESCENDING, fp32 sequential; the dense zeros contribute
 * +0.0f terms that do not change an fp32 running sum.
 * order=1: fp64 accumulation (the "exact" variant used at sizes where sequential fp32 is itself
 * less accurate than 1e-5).
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_spmm(int32_t N, const int64_t *ptr, const int32_t *idx, const float *val, const float *P,
                      int64_t ldp, int32_t F, float *Y, int64_t ldy, int order) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t r = 0; r < N; r++) {
        if (order == 0) {
            for (int32_t f = 0; f < F; f++) Y[(int64_t)r * ldy + f] = 0.0f;
            for (int64_t k = ptr[r + 1] - 1; k >= ptr[r]; k--) {
                const float a = val[k];
                const float *p = P + (int64_t)idx[k] * ldp;
                float *y = Y + (int64_t)r * ldy;
                for (int32_t f = 0; f < F; f++) { float t = p[f] * a; y[f] = y[f] + t; }
            }
        } else {
            double acc[1024];
            for (int32_t f0 = 0; f0 < F; f0 += 1024) {
                int32_t fn = F - f0 < 1024 ? F - f0 : 1024;
                for (int32_t f = 0; f < fn; f++) acc[f] = 0.0;
                for (int64_t k = ptr[r]; k < ptr[r + 1]; k++) {
                    const double a = val[k];
                    const float *p = P + (int64_t)idx[k] * ldp + f0;
                    for (int32_t f = 0; f < fn; f++) acc[f] += a * (double)p[f];
                }
                for (int32_t f = 0; f < fn; f++) Y[(int64_t)r * ldy + f0 + f] = (float)acc[f];
            }
        }
    }
}


/* ---- order = 1 ("exact") fast path ---------------------------------------------------------------
 * fp64 accumulation of fp32 x fp32 products (each product is exact in fp64, so the result does not depend on
 * the summation order beyond 1e-16): used as the checker at sizes where the reference's own sequential fp32
 * sums are less accurate than the 1e-5 contract.  Because the order is free here, this path is register-blocked
 * (4 x 12 fp64 accumulators, GCC vector extensions) so that a products-shaped step takes seconds, not minutes.
 * Nothing below is used by order 0, which stays the literal reference order. */
typedef double orc_v4d __attribute__((vector_size(32)));
typedef float orc_v4f __attribute__((vector_size(16)));
static inline orc_v4d orc_ld4(const float *p) {
    orc_v4f f;
    memcpy(&f, p, 16);
    return __builtin_convertvector(f, orc_v4d);
}
/* out[m][n] (m < 4, n < 12; leading dimension 12) = sum_{k<K} A[m*sam + k*sak] * Bp[k*ldb + n]; rows m >= mr repeat row mr-1 */
static void orc_mk_4x12(const float *A, int64_t sam, int64_t sak, int mr, const float *Bp, int64_t ldb, int64_t K,
                        double *out) {
    const float *a0 = A, *a1 = A + (mr > 1 ? 1 : 0) * sam, *a2 = A + (mr > 2 ? 2 : mr - 1) * sam,
                *a3 = A + (mr > 3 ? 3 : mr - 1) * sam;
    orc_v4d c00 = {0, 0, 0, 0}, c01 = c00, c02 = c00, c10 = c00, c11 = c00, c12 = c00, c20 = c00, c21 = c00, c22 = c00,
            c30 = c00, c31 = c00, c32 = c00;
    for (int64_t k = 0; k < K; k++) {
        const float *b = Bp + k * ldb;
        const orc_v4d b0 = orc_ld4(b), b1 = orc_ld4(b + 4), b2 = orc_ld4(b + 8);
        const double x0 = a0[k * sak], x1 = a1[k * sak], x2 = a2[k * sak], x3 = a3[k * sak];
        const orc_v4d v0 = {x0, x0, x0, x0}, v1 = {x1, x1, x1, x1}, v2 = {x2, x2, x2, x2}, v3 = {x3, x3, x3, x3};
        c00 += v0 * b0; c01 += v0 * b1; c02 += v0 * b2;
        c10 += v1 * b0; c11 += v1 * b1; c12 += v1 * b2;
        c20 += v2 * b0; c21 += v2 * b1; c22 += v2 * b2;
        c30 += v3 * b0; c31 += v3 * b1; c32 += v3 * b2;
    }
    memcpy(out + 0, &c00, 32);  memcpy(out + 4, &c01, 32);  memcpy(out + 8, &c02, 32);
    memcpy(out + 12, &c10, 32); memcpy(out + 16, &c11, 32); memcpy(out + 20, &c12, 32);
    memcpy(out + 24, &c20, 32); memcpy(out + 28, &c21, 32); memcpy(out + 32, &c22, 32);
    memcpy(out + 36, &c30, 32); memcpy(out + 40, &c31, 32); memcpy(out + 44, &c32, 32);
}
/* C[M,N] (fp32) = A[M,K] * Bp[K,Np] with Bp packed, zero padded to Np = multiple of 12 columns */
static void orc_gemm_rows_f64(int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *Bp, int32_t Np,
                              float *C, int64_t ldc) {
#pragma omp parallel for schedule(static)
    for (int64_t i0 = 0; i0 < M; i0 += 4) {
        const int mr = M - i0 < 4 ? (int)(M - i0) : 4;
        double t[48];
        for (int32_t j0 = 0; j0 < Np; j0 += 12) {
            orc_mk_4x12(A + i0 * lda, lda, 1, mr, Bp + j0, Np, K, t);
            for (int m = 0; m < mr; m++)
                for (int32_t j = j0; j < j0 + 12 && j < N; j++) C[(i0 + m) * ldc + j] = (float)t[m * 12 + (j - j0)];
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Dense feature transforms (nn::Linear, src/nn.cpp:205-211; MatMul::_backward, operation.h:504-534;
 * Transpose::_backward, operation.h:416-433).  All row-major.
 *   NT: C[M,N] = A[M,K] * B[N,K]^T     forward  P = H * W^T
 *   NN: C[M,N] = A[M,K] * B[K,N]       backward dH = dP * W
 *   TN: C[K1,K2] = A[M,K1]^T * B[M,K2] backward dW = dP^T * H   (reduction over the M rows)
 * order 0 = reference order (descending reduction index, fp32, no FMA); order 1 = fp64 accumulate.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_gemm_nt(int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B, int64_t ldb,
                         float *C, int64_t ldc, int order) {
    if (order != 0) { /* pack B^T: Bp[k][j] = B[j][k] */
        const int32_t Np = (N + 11) / 12 * 12;
        float *Bp = (float *)calloc((size_t)K * Np, 4);
        for (int32_t j = 0; j < N; j++)
            for (int32_t k = 0; k < K; k++) Bp[(size_t)k * Np + j] = B[(int64_t)j * ldb + k];
        orc_gemm_rows_f64(M, N, K, A, lda, Bp, Np, C, ldc);
        free(Bp);
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; i++)
        for (int32_t j = 0; j < N; j++) {
            const float *a = A + i * lda, *b = B + (int64_t)j * ldb;
            if (order == 0) {
                float s = 0.0f;
                for (int32_t k = K - 1; k >= 0; k--) { float t = b[k] * a[k]; s = s + t; }
                C[i * ldc + j] = s;
            } else {
                double s = 0.0;
                for (int32_t k = 0; k < K; k++) s += (double)a[k] * (double)b[k];
                C[i * ldc + j] = (float)s;
            }
        }
}
ORC_API void orc_gemm_nn(int64_t M, int32_t N, int32_t K, const float *A, int64_t lda, const float *B, int64_t ldb,
                         float *C, int64_t ldc, int order) {
    if (order != 0) {
        const int32_t Np = (N + 11) / 12 * 12;
        float *Bp = (float *)calloc((size_t)K * Np, 4);
        for (int32_t k = 0; k < K; k++) memcpy(Bp + (size_t)k * Np, B + (int64_t)k * ldb, (size_t)N * 4);
        orc_gemm_rows_f64(M, N, K, A, lda, Bp, Np, C, ldc);
        free(Bp);
        return;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < M; i++) {
        const float *a = A + i * lda;
        if (order == 0) {
            for (int32_t j = 0; j < N; j++) {
                float s = 0.0f;
                for (int32_t k = K - 1; k >= 0; k--) { float t = B[(int64_t)k * ldb + j] * a[k]; s = s + t; }
                C[i * ldc + j] = s;
            }
        } else {
            double acc[2048];
            for (int32_t j = 0; j < N; j++) acc[j] = 0.0;
            for (int32_t k = 0; k < K; k++) {
                const double av = a[k];
                const float *b = B + (int64_t)k * ldb;
                for (int32_t j = 0; j < N; j++) acc[j] += av * (double)b[j];
            }
            for (int32_t j = 0; j < N; j++) C[i * ldc + j] = (float)acc[j];
        }
    }
}
ORC_API void orc_gemm_tn(int64_t M, int32_t K1, int32_t K2, const float *A, int64_t lda, const float *B, int64_t ldb,
                         float *C, int64_t ldc, int order) {
    if (order == 0) {
        /* out[k1][k2] = sum_{i desc} B[i][k2]*A[i][k1]; running sums updated row by row from the
         * last row to the first give exactly that order for every output at once. */
        for (int32_t a = 0; a < K1; a++)
            for (int32_t b = 0; b < K2; b++) C[(int64_t)a * ldc + b] = 0.0f;
#pragma omp parallel for schedule(static)
        for (int32_t a = 0; a < K1; a++) {
            float *c = C + (int64_t)a * ldc;
            for (int64_t i = M - 1; i >= 0; i--) {
                const float av = A[i * lda + a];
                const float *b = B + i * ldb;
                for (int32_t j = 0; j < K2; j++) { float t = b[j] * av; c[j] = c[j] + t; }
            }
        }
    } else {
        /* node rows are cut into blocks of TNB; a thread packs the block of B (zero padded to 12-column tiles) and adds
         * the block's 4 x 12 tile products into its private fp64 copy of the output; copies are summed at the end */
        enum { TNB = 256 };
        const int32_t Np = (K2 + 11) / 12 * 12;
        double *acc = (double *)calloc((size_t)K1 * K2, sizeof(double));
#pragma omp parallel
        {
            double *loc = (double *)calloc((size_t)K1 * Np, sizeof(double));
            float *Bp = (float *)calloc((size_t)TNB * Np, 4);
            double t[48];
#pragma omp for schedule(static)
            for (int64_t i0 = 0; i0 < M; i0 += TNB) {
                const int64_t nb = M - i0 < TNB ? M - i0 : TNB;
                for (int64_t i = 0; i < nb; i++) memcpy(Bp + (size_t)i * Np, B + (i0 + i) * ldb, (size_t)K2 * 4);
                for (int32_t x0 = 0; x0 < K1; x0 += 4) {
                    const int mr = K1 - x0 < 4 ? K1 - x0 : 4;
                    for (int32_t j0 = 0; j0 < Np; j0 += 12) {
                        orc_mk_4x12(A + i0 * lda + x0, 1, lda, mr, Bp + j0, Np, nb, t);
                        for (int m = 0; m < mr; m++)
                            for (int j = 0; j < 12; j++) loc[(size_t)(x0 + m) * Np + j0 + j] += t[m * 12 + j];
                    }
                }
            }
#pragma omp critical
            for (int32_t x = 0; x < K1; x++)
                for (int32_t j = 0; j < K2; j++) acc[(size_t)x * K2 + j] += loc[(size_t)x * Np + j];
            free(loc); free(Bp);
        }
        for (int32_t x = 0; x < K1; x++)
            for (int32_t j = 0; j < K2; j++) C[(int64_t)x * ldc + j] = (float)acc[(size_t)x * K2 + j];
        free(acc);
    }
}

/* Bias add (Add, operation.h:102-129 / functional.h:162-187) and ReLU (nn.cpp:229-237 ->
 * functional::mask, functional.h:443-471: out = x>0 ? x : 0, NaN -> 0). H may alias Z or be NULL. */
ORC_API void orc_bias_relu(int64_t N, int32_t F, const float *Y, int64_t ldy, const float *bias, float *Z, int64_t ldz,
                           float *H, int64_t ldh) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++)
        for (int32_t f = 0; f < F; f++) {
            float z = Y[i * ldy + f] + (bias ? bias[f] : 0.0f);
            if (Z) Z[i * ldz + f] = z;
            if (H) H[i * ldh + f] = z > 0.0f ? z : 0.0f;
        }
}
/* ReLU mask backward (Mask::_backward, operation.h:557-562: grad zeroed where cond<=0) */
ORC_API void orc_relu_bwd(int64_t N, int32_t F, const float *dH, int64_t ldd, const float *Z, int64_t ldz, float *dZ,
                          int64_t ldo) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++)
        for (int32_t f = 0; f < F; f++) dZ[i * ldo + f] = Z[i * ldz + f] > 0.0f ? dH[i * ldd + f] : 0.0f;
}
/* Bias gradient: db[f] = sum over rows ASCENDING (Add::_backward -> sum_to_size -> sum(0),
 * tensor.h:618-638, functional.h:285-288). */
ORC_API void orc_bias_grad(int64_t N, int32_t F, const float *dZ, int64_t ldd, float *db, int order) {
    if (order != 0) { /* fp64 column sums, one pass over the matrix (rows split over the threads) */
        double *acc = (double *)calloc((size_t)F, sizeof(double));
#pragma omp parallel
        {
            double *loc = (double *)calloc((size_t)F, sizeof(double));
#pragma omp for schedule(static)
            for (int64_t i = 0; i < N; i++)
                for (int32_t f = 0; f < F; f++) loc[f] += (double)dZ[i * ldd + f];
#pragma omp critical
            for (int32_t f = 0; f < F; f++) acc[f] += loc[f];
            free(loc);
        }
        for (int32_t f = 0; f < F; f++) db[f] = (float)acc[f];
        free(acc);
        return;
    }
#pragma omp parallel for schedule(static)
    for (int32_t f = 0; f < F; f++) {
        if (order == 0) {
            float s = 0.0f;
            for (int64_t i = 0; i < N; i++) s = s + dZ[i * ldd + f];
            db[f] = s;
        } else {
            double s = 0.0;
            for (int64_t i = 0; i < N; i++) s += (double)dZ[i * ldd + f];
            db[f] = (float)s;
        }
    }
}

/* Loss: nn::cross_entropy_loss (src/nn.cpp:442-453):
 *   L = (1/N) * sum_i -log( exp(z[i,y_i]) / (sum_c exp(z[i,c]) + 1e-20) )     no max-shift
 * dZ (optional) = (softmax(Z) - onehot(y)) / N with softmax = exp(z - log(sum exp z))
 * (nn::softmax, src/nn.cpp:270-278). Returns the loss. */
ORC_API float orc_softmax_xent(int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y, float *dZ,
                               int64_t ldd, int order) {
    float *li = (float *)malloc((size_t)N * 4);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; i++) {
        const float *z = Z + i * ldz;
        float s = 0.0f;
        for (int32_t c = 0; c < C; c++) s = s + expf(z[c]);
        float num = expf(z[y[i]]);
        float q = num / (s + 1e-20f);
        li[i] = -logf(q);
        if (dZ) {
            float ls = logf(s);
            for (int32_t c = 0; c < C; c++) {
                float sm = expf(z[c] - ls);
                if (c == y[i]) sm -= 1.0f;
                dZ[i * ldd + c] = sm / (float)N;
            }
        }
    }
    float loss;
    if (order == 0) {
        float s = 0.0f;
        for (int64_t i = 0; i < N; i++) s = s + li[i];
        loss = s / (float)N;
    } else {
        double s = 0.0;
        for (int64_t i = 0; i < N; i++) s += (double)li[i];
        loss = (float)(s / (double)N);
    }
    free(li);
    return loss;
}

/* SGD with torch.optim.SGD semantics (the documented intent of nn::SGD, include/nn.h:165-178; the
 * reference body, src/nn.cpp:395-417, is broken — bug B4).  `first` = no momentum buffer yet. */
ORC_API void orc_sgd_step(int64_t n, float *p, const float *g, float *vel, float lr, float momentum, float dampening,
                          float weight_decay, int nesterov, int first) {
    for (int64_t i = 0; i < n; i++) {
        float d = g[i];
        if (weight_decay != 0.0f) d = d + weight_decay * p[i];
        if (momentum != 0.0f) {
            float v = first ? d : momentum * vel[i] + (1.0f - dampening) * d;
            vel[i] = v;
            d = nesterov ? d + momentum * v : v;
        }
        p[i] = p[i] - lr * d;
    }
}

/* ------------------------------------------------------------------------------------------------
 * BatchNorm over the node dimension, training statistics (nn::BatchNorm::forward, src/nn.cpp:301-330):
 *   mean_c = sum_r x[r,c] / N                        x->mean(-2,true): ascending fp32 sum (functional.h:266-307)
 *   var_c  = sum_r (x[r,c]-mean_c)^2 / max(0, N-0)   x->var(-2, 0, true): two passes, std::pow(.,2) (functional.h:383-387),
 *                                                    summed in DESCENDING row order (libstdc++ _Expr::sum)
 *   y      = (x - mean) / pow(var + eps, 0.5) * gamma + beta     (nn.cpp:314-318), optional ReLU (nn.cpp:229-237)
 * order=1: fp64 statistics.  mean/var are returned for the backward.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_batchnorm_fwd(int64_t N, int32_t F, const float *X, int64_t ldx, const float *gamma, const float *beta,
                               float eps, int relu, float *Y, int64_t ldy, float *mean, float *var, int order) {
    for (int32_t c = 0; c < F; c++) {
        if (order == 0) {
            float s = 0.0f;
            for (int64_t r = 0; r < N; r++) s = s + X[r * ldx + c];
            const float m = s / (float)N;
            /* std::pow(expr, 2).sum() is libstdc++'s _Expr::sum(): it starts from the LAST element and walks down */
            float q = 0.0f;
            for (int64_t r = N - 1; r >= 0; r--) { float d = X[r * ldx + c] - m; float t = powf(d, 2.0f); q = (r == N - 1) ? t : q + t; }
            mean[c] = m;
            var[c] = q / (float)(N > 0 ? N : 0);
        } else {
            double s = 0.0;
            for (int64_t r = 0; r < N; r++) s += X[r * ldx + c];
            const double m = s / (double)N;
            double q = 0.0;
            for (int64_t r = 0; r < N; r++) { double d = X[r * ldx + c] - m; q += d * d; }
            mean[c] = (float)m;
            var[c] = (float)(q / (double)N);
        }
    }
    for (int64_t r = 0; r < N; r++)
        for (int32_t c = 0; c < F; c++) {
            float v = (X[r * ldx + c] - mean[c]) / powf(var[c] + eps, 0.5f);
            v = v * gamma[c];
            if (beta) v = v + beta[c];
            if (relu) v = v > 0.0f ? v : 0.0f;
            Y[r * ldy + c] = v;
        }
}

/* Gradient of y = relu?(BN(x)) w.r.t. x, gamma, beta — the standard batch-norm backward (the reference's autograd
 * loses the fan-out contributions of x here, bug B2: parity unpinned in the reference, pinned against torch autograd in
 * tests/test_oracle_vs_reference.py).  Yout = the forward output (ReLU mask), may be NULL when relu == 0. */
ORC_API void orc_batchnorm_bwd(int64_t N, int32_t F, const float *X, int64_t ldx, const float *mean, const float *var,
                               const float *gamma, float eps, int relu, const float *Yout, int64_t ldy, const float *dY,
                               int64_t ldd, float *dX, int64_t ldo, float *dgamma, float *dbeta) {
    for (int32_t c = 0; c < F; c++) {
        const double istd = 1.0 / sqrt((double)var[c] + (double)eps);
        double sg = 0.0, sgx = 0.0;
        for (int64_t r = 0; r < N; r++) {
            double g = dY[r * ldd + c];
            if (relu && !(Yout[r * ldy + c] > 0.0f)) g = 0.0;
            const double xh = ((double)X[r * ldx + c] - (double)mean[c]) * istd;
            sg += g;
            sgx += g * xh;
        }
        if (dbeta) dbeta[c] = (float)sg;
        if (dgamma) dgamma[c] = (float)sgx;
        for (int64_t r = 0; r < N; r++) {
            double g = dY[r * ldd + c];
            if (relu && !(Yout[r * ldy + c] > 0.0f)) g = 0.0;
            const double xh = ((double)X[r * ldx + c] - (double)mean[c]) * istd;
            dX[r * ldo + c] = (float)((double)gamma[c] * istd * (g - sg / (double)N - xh * sgx / (double)N));
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * nn::LayerNorm (src/nn.cpp:332-353): per row  y = (x - mean) / pow(var + eps, 0.5) * gamma + beta,
 *   mean = x->mean(-1,true) (ascending fp32 sum / F), var = x->var(-1, 0, true) (two passes, correction 0);
 * optional ReLU (nn::MLP puts nn::ReLU right after, include/nn.h:193-214).  order=1: fp64 statistics.
 * mean / rstd (= 1/sqrt(var+eps)) per row are returned for the backward.
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_layernorm_fwd(int64_t N, int32_t F, const float *X, int64_t ldx, const float *gamma, const float *beta,
                               float eps, int relu, float *Y, int64_t ldy, float *mean, float *rstd, int order) {
    for (int64_t r = 0; r < N; r++) {
        const float *x = X + r * ldx;
        float m, v;
        if (order == 0) {
            float s = 0.0f;
            for (int32_t c = 0; c < F; c++) s = s + x[c];
            m = s / (float)F;
            float q = 0.0f; /* _Expr::sum(): descending */
            for (int32_t c = F - 1; c >= 0; c--) { float d = x[c] - m; float t = powf(d, 2.0f); q = (c == F - 1) ? t : q + t; }
            v = q / (float)F;
        } else {
            double s = 0.0;
            for (int32_t c = 0; c < F; c++) s += x[c];
            double md = s / F, q = 0.0;
            for (int32_t c = 0; c < F; c++) { double d = x[c] - md; q += d * d; }
            m = (float)md;
            v = (float)(q / F);
        }
        mean[r] = m;
        rstd[r] = 1.0f / powf(v + eps, 0.5f);
        for (int32_t c = 0; c < F; c++) {
            float o = (x[c] - m) / powf(v + eps, 0.5f);
            if (gamma) o = o * gamma[c];
            if (beta) o = o + beta[c];
            if (relu) o = o > 0.0f ? o : 0.0f;
            Y[r * ldy + c] = o;
        }
    }
}
/* standard layer-norm gradient (the reference autograd loses the fan-out terms, bug B2; pinned against torch) */
ORC_API void orc_layernorm_bwd(int64_t N, int32_t F, const float *X, int64_t ldx, const float *mean, const float *rstd,
                               const float *gamma, int relu, const float *Yout, int64_t ldy, const float *dY, int64_t ldd,
                               float *dX, int64_t ldo, float *dgamma, float *dbeta) {
    double *sg = (double *)calloc((size_t)F, 8), *sgx = (double *)calloc((size_t)F, 8);
    for (int64_t r = 0; r < N; r++) {
        double a = 0.0, b = 0.0;
        for (int32_t c = 0; c < F; c++) {
            double g = dY[r * ldd + c];
            if (relu && !(Yout[r * ldy + c] > 0.0f)) g = 0.0;
            const double xh = ((double)X[r * ldx + c] - mean[r]) * rstd[r];
            sg[c] += g;
            sgx[c] += g * xh;
            const double gg = g * (gamma ? gamma[c] : 1.0);
            a += gg;
            b += gg * xh;
        }
        for (int32_t c = 0; c < F; c++) {
            double g = dY[r * ldd + c];
            if (relu && !(Yout[r * ldy + c] > 0.0f)) g = 0.0;
            const double xh = ((double)X[r * ldx + c] - mean[r]) * rstd[r];
            const double gg = g * (gamma ? gamma[c] : 1.0);
            dX[r * ldo + c] = (float)(rstd[r] * (gg - a / F - xh * b / F));
        }
    }
    for (int32_t c = 0; c < F; c++) { if (dgamma) dgamma[c] = (float)sgx[c]; if (dbeta) dbeta[c] = (float)sg[c]; }
    free(sg); free(sgx);
}
/* nn::tanh as written (src/nn.cpp:355-364): s = x + 1e-12; (exp(s) - exp(-s)) / (exp(s) + exp(-s)) in fp32 */
ORC_API void orc_tanh_fwd(int64_t n, const float *x, float *y) {
    for (int64_t i = 0; i < n; i++) {
        const float s = x[i] + 1e-12f;
        const float ep = expf(s), em = expf(-s);
        y[i] = (ep - em) / (ep + em);
    }
}
/* Dropout keep-mask of the build (the reference seeds a fresh engine from time(), bug B6: unpinned): element i is kept
 * when u_i = (splitmix-hash(seed, 77, i) >> 40) * 2^-24 >= p; kept values are scaled by 1/(1-p) (src/nn.cpp:246-266). */
ORC_API void orc_dropout_fwd(int64_t n, const float *x, float p, uint64_t seed, float *y) {
    const float scale = 1.0f / (1.0f - p);
    for (int64_t i = 0; i < n; i++) {
        const float u = (float)(orc_hash3(seed, 77, (uint64_t)i) >> 40) * (1.0f / 16777216.0f);
        y[i] = u >= p ? x[i] * scale : 0.0f;
    }
}

/* The factorised normalisation of graph::GCNConv::forward as written (src/graph.cpp:176-185) on the loop-free
 * adjacency A0 (CSR without diagonal):  deg = rowsum(A0)+1, dinv = pow(deg,-0.5), norm = (A0 dinv) * dinv, where the
 * matrix-vector product is the reference's descending-k fp32 dot product. */
ORC_API void orc_aswritten_norm(int32_t N, const int64_t *rowptr, const int32_t *colidx, float *dinv, float *norm) {
    for (int32_t r = 0; r < N; r++) dinv[r] = powf((float)(rowptr[r + 1] - rowptr[r]) + 1.0f, -0.5f);
    for (int32_t r = 0; r < N; r++) {
        float s = 0.0f;
        for (int64_t k = rowptr[r + 1] - 1; k >= rowptr[r]; k--) { float t = dinv[colidx[k]] * 1.0f; s = s + t; }
        norm[r] = s * dinv[r];
    }
}

/* Adam with torch.optim.Adam semantics — the documented intent of nn::Adam (include/nn.h:180-188); the reference
 * body (src/nn.cpp:419-441) divides by sqrt(v)*eps and uses the parameter index as step: parity unpinned in the
 * reference, pinned against torch in tests/test_oracle_vs_reference.py.  step counts from 1. */
ORC_API void orc_adam_step(int64_t n, float *p, const float *g, float *m, float *v, float lr, float b1, float b2,
                           float eps, float weight_decay, int64_t step) {
    const float bc1 = (float)(1.0 - pow((double)b1, (double)step)), bc2 = (float)(1.0 - pow((double)b2, (double)step));
    for (int64_t i = 0; i < n; i++) {
        float d = g[i];
        if (weight_decay != 0.0f) d = d + weight_decay * p[i];
        m[i] = b1 * m[i] + (1.0f - b1) * d;
        v[i] = b2 * v[i] + (1.0f - b2) * d * d;
        const float denom = sqrtf(v[i]) / sqrtf(bc2) + eps;
        p[i] = p[i] - (lr / bc1) * (m[i] / denom);
    }
}

/* Loss over the rows selected by a node mask (Data::set_mask, src/graph.cpp:130-151): the reference would slice
 * logits/targets by the mask and call cross_entropy_loss (src/nn.cpp:442-453) on the slice: mean over the selected
 * rows of -log(exp(z_y)/(sum exp(z)+1e-20)), ascending row order.  dZ (optional): (softmax - onehot)/n_sel on selected
 * rows, 0 elsewhere.  Returns the loss; *n_sel_out = number of selected rows. */
ORC_API float orc_softmax_xent_masked(int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y,
                                      const uint8_t *mask, float *dZ, int64_t ldd, int64_t *n_sel_out) {
    int64_t n_sel = 0;
    for (int64_t r = 0; r < N; r++) n_sel += mask[r] ? 1 : 0;
    if (n_sel_out) *n_sel_out = n_sel;
    double acc = 0.0;
    const float inv_n = 1.0f / (float)n_sel;
    for (int64_t r = 0; r < N; r++) {
        if (!mask[r]) {
            if (dZ) for (int32_t c = 0; c < C; c++) dZ[r * ldd + c] = 0.0f;
            continue;
        }
        const float *z = Z + r * ldz;
        double s = 0.0;
        for (int32_t c = 0; c < C; c++) s += exp((double)z[c]);
        acc += -log(exp((double)z[y[r]]) / (s + 1e-20));
        if (dZ)
            for (int32_t c = 0; c < C; c++) {
                float pr = (float)(exp((double)z[c]) / s);
                if (c == y[r]) pr -= 1.0f;
                dZ[r * ldd + c] = pr * inv_n;
            }
    }
    return (float)(acc / (double)n_sel);
}

/* rows (selected by mask, all when NULL) whose first-maximum column equals the label (tensor::argmax,
 * include/tensor.h:645-648 uses std::max_element: first maximum) */
ORC_API int64_t orc_argmax_correct(int64_t N, int32_t C, const float *Z, int64_t ldz, const int32_t *y, const uint8_t *mask) {
    int64_t n = 0;
    for (int64_t r = 0; r < N; r++) {
        if (mask && !mask[r]) continue;
        const float *z = Z + r * ldz;
        int32_t bi = 0;
        for (int32_t c = 1; c < C; c++)
            if (z[c] > z[bi]) bi = c;
        n += bi == y[r];
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------
 * 1-D row partition.  No counterpart in the reference.
 *   part_ptr[p] = min(N, p * ceil(N/P)),  p = 0..P
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_partition_ptr(int64_t N, int32_t P, int64_t *part_ptr) {
    int64_t chunk = (N + P - 1) / P;
    for (int32_t p = 0; p <= P; p++) {
        int64_t v = (int64_t)p * chunk;
        part_ptr[p] = v < N ? v : N;
    }
}
/* Local CSR block of rank p: rows [lo,hi) of the global CSR with rowptr rebased to 0 (column ids stay
 * global: the halo exchange materialises the full feature matrix in global row order). */
ORC_API int64_t orc_partition_rows(const int64_t *rowptr, const int32_t *colidx, const float *val, int64_t lo,
                                   int64_t hi, int64_t *l_rowptr, int32_t *l_colidx, float *l_val) {
    int64_t base = rowptr[lo];
    for (int64_t r = lo; r <= hi; r++) l_rowptr[r - lo] = rowptr[r] - base;
    int64_t n = rowptr[hi] - base;
    memcpy(l_colidx, colidx + base, (size_t)n * 4);
    if (val && l_val) memcpy(l_val, val + base, (size_t)n * 4);
    return n;
}
/* interior flag per local row: 1 when every column lies in [lo,hi) (no halo needed) */
ORC_API int64_t orc_partition_interior(const int64_t *l_rowptr, const int32_t *l_colidx, int64_t nrows, int64_t lo,
                                       int64_t hi, uint8_t *interior) {
    int64_t cnt = 0;
    for (int64_t r = 0; r < nrows; r++) {
        uint8_t in = 1;
        for (int64_t k = l_rowptr[r]; k < l_rowptr[r + 1]; k++)
            if (l_colidx[k] < lo || l_colidx[k] >= hi) { in = 0; break; }
        interior[r] = in;
        cnt += in;
    }
    return cnt;
}

/* Halo of the local block rows [lo,hi): halo_ids = sorted unique column ids outside [lo,hi) (capacity: n_cols), and the
 * locally renumbered column array: an owned column c becomes c - lo, a halo column n_loc + its position in halo_ids
 * (so the aggregation input is one matrix [own rows ; halo rows]).  Returns the halo size. */
ORC_API int64_t orc_partition_halo(const int64_t *l_rowptr, const int32_t *l_colidx, int64_t nrows, int64_t n_cols,
                                   int64_t lo, int64_t hi, int32_t *halo_ids, int32_t *local_colidx) {
    int32_t *pos = (int32_t *)malloc((size_t)(n_cols > 0 ? n_cols : 1) * 4);
    for (int64_t c = 0; c < n_cols; c++) pos[c] = -1;
    const int64_t nnz = l_rowptr[nrows];
    for (int64_t k = 0; k < nnz; k++) {
        int32_t c = l_colidx[k];
        if (c < lo || c >= hi) pos[c] = 0;
    }
    int64_t n_halo = 0;
    for (int64_t c = 0; c < n_cols; c++)
        if (pos[c] == 0) { halo_ids[n_halo] = (int32_t)c; pos[c] = (int32_t)n_halo++; }
    if (local_colidx)
        for (int64_t k = 0; k < nnz; k++) {
            int32_t c = l_colidx[k];
            local_colidx[k] = (c >= lo && c < hi) ? (int32_t)(c - lo) : (int32_t)(hi - lo) + pos[c];
        }
    free(pos);
    return n_halo;
}

/* ------------------------------------------------------------------------------------------------
 * Whole train step (forward + loss + backward + SGD) for an L-layer GCN in the reference's operation
 * order:  Z_l = Ahat (H_{l-1} W_l^T) + b_l,  H_l = ReLU(Z_l) for l < L;  logits = Z_L.
 * Outputs (any may be NULL): Zs[l] (N x F_l), dWs[l], dbs[l], dZL. Weights updated in place when lr>0
 * (plain SGD).  CSC arrays give Ahat^T for the backward aggregation (operation.h:524-531).
 * ---------------------------------------------------------------------------------------------- */
ORC_API float orc_gcn_train_step(int32_t N, int32_t L, const int32_t *dims, const int64_t *rowptr,
                                 const int32_t *colidx, const float *val, const int64_t *colptr, const int32_t *rowidx,
                                 const float *valT, const float *X, const int32_t *y, float **W, float **b, float **Zs,
                                 float **dWs, float **dbs, float *dZL, float lr, int order) {
    float **H = (float **)calloc((size_t)L + 1, sizeof(float *));
    float **Z = (float **)calloc((size_t)L + 1, sizeof(float *));
    H[0] = (float *)X;
    for (int32_t l = 1; l <= L; l++) {
        int32_t Fi = dims[l - 1], Fo = dims[l];
        float *P = (float *)malloc((size_t)N * Fo * 4);
        float *Y = (float *)malloc((size_t)N * Fo * 4);
        Z[l] = (float *)malloc((size_t)N * Fo * 4);
        orc_gemm_nt(N, Fo, Fi, H[l - 1], Fi, W[l - 1], Fi, P, Fo, order);
        orc_spmm(N, rowptr, colidx, val, P, Fo, Fo, Y, Fo, order);
        if (l < L) {
            H[l] = (float *)malloc((size_t)N * Fo * 4);
            orc_bias_relu(N, Fo, Y, Fo, b[l - 1], Z[l], Fo, H[l], Fo);
        } else {
            orc_bias_relu(N, Fo, Y, Fo, b[l - 1], Z[l], Fo, NULL, 0);
        }
        if (Zs && Zs[l - 1]) memcpy(Zs[l - 1], Z[l], (size_t)N * Fo * 4);
        free(P); free(Y);
    }
    int32_t C = dims[L];
    float *dZ = (float *)malloc((size_t)N * C * 4);
    float loss = orc_softmax_xent(N, C, Z[L], C, y, dZ, C, order);
    if (dZL) memcpy(dZL, dZ, (size_t)N * C * 4);
    for (int32_t l = L; l >= 1; l--) {
        int32_t Fi = dims[l - 1], Fo = dims[l];
        float *db = (float *)malloc((size_t)Fo * 4);
        float *dW = (float *)malloc((size_t)Fo * Fi * 4);
        float *dP = (float *)malloc((size_t)N * Fo * 4);
        orc_bias_grad(N, Fo, dZ, Fo, db, order);
        orc_spmm(N, colptr, rowidx, valT, dZ, Fo, Fo, dP, Fo, order);
        /* dW = (H^T dP)^T : out[fo][fi] = sum_i dP[i][fo] * H[i][fi] */
        orc_gemm_tn(N, Fo, Fi, dP, Fo, H[l - 1], Fi, dW, Fi, order);
        float *dHprev = NULL;
        if (l > 1) {
            dHprev = (float *)malloc((size_t)N * Fi * 4);
            orc_gemm_nn(N, Fi, Fo, dP, Fo, W[l - 1], Fi, dHprev, Fi, order);
        }
        if (dWs && dWs[l - 1]) memcpy(dWs[l - 1], dW, (size_t)Fo * Fi * 4);
        if (dbs && dbs[l - 1]) memcpy(dbs[l - 1], db, (size_t)Fo * 4);
        if (lr > 0.0f) {
            orc_sgd_step((int64_t)Fo * Fi, W[l - 1], dW, NULL, lr, 0, 0, 0, 0, 1);
            orc_sgd_step(Fo, b[l - 1], db, NULL, lr, 0, 0, 0, 0, 1);
        }
        free(db); free(dW); free(dP); free(dZ);
        dZ = NULL;
        if (l > 1) {
            dZ = (float *)malloc((size_t)N * Fi * 4);
            orc_relu_bwd(N, Fi, dHprev, Fi, Z[l - 1], Fi, dZ, Fi);
            free(dHprev);
        }
    }
    for (int32_t l = 1; l <= L; l++) { free(Z[l]); if (l < L) free(H[l]); }
    free(H); free(Z);
    return loss;
}
