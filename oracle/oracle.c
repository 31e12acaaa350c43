/*
 * oracle.c — CPU restatement of reductive 0.9.0's product-quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  It is the parity authority the CUDA path is
 * checked against, and the timed CPU baseline of bench.py.  It is never on the product path.
 *
 * What is restated, with the reference lines each function follows:
 *   src/linalg.rs:150-180   squared Euclidean distance matrix  (orc_sqdist_batch)
 *   src/linalg.rs:118-148   vector-vs-matrix distances         (orc_sqdist_vec)
 *   src/kmeans.rs:111-159   nearest-centroid argmin            (orc_cluster_assignment[s])
 *   src/kmeans.rs:166-198   centroid update                    (orc_update_centroids)
 *   src/kmeans.rs:263-360   Lloyd loop + mean squared error    (orc_kmeans_*)
 *   src/pq/primitives.rs    quantize / reconstruct loops       (orc_quantize_*, orc_reconstruct*)
 *   src/pq/pq.rs:63-100,144-249,256-347   validation, training, projection before/after
 *
 * Third-party arithmetic the reference bottoms out in is NOT in /root/reference (Cargo.toml:13
 * `ndarray = "0.15"`, transitively `matrixmultiply` 0.3.x, no Cargo.lock), and the reference
 * cannot be compiled in this image (no rustc/cargo).  Their published algorithms are restated:
 *   - ndarray numeric_util::unrolled_dot: 8 independent accumulators, separate multiply and add
 *     (rustc never contracts), combined ((((0+(p0+p4))+(p1+p5))+(p2+p6))+(p3+p7)), then the <8
 *     tail added sequentially.  Non-contiguous 1-D dot: plain sequential sum.
 *   - matrixmultiply sgemm on an FMA-capable x86-64 host: every C element owns one accumulator
 *     updated by a fused multiply-add sequentially over k, K split in blocks of kc = 256 whose
 *     partial products are combined with a plain add (first block: C = AB, later: C = C + AB).
 *   - ndarray mat-vec without BLAS: one row.dot(x) per row.
 * PARITY PINNING: the oracle reproduces every known-answer vector in the reference's own tests
 * for this path (pq.rs:378-490, kmeans.rs:380-435,504-519, linalg.rs:291-313; tests/test_oracle_golden.py).
 * Those vectors are exactly representable, so they pin semantics (argmin, update, mse, gather) but
 * not the accumulation order: for tie-breaking, NaN order and last-bit rounding at realistic shapes
 * parity is UNPINNED by anything the reference ships; this file is then the stated authority.
 *
 * Build: see oracle/Makefile (-O2 -mavx2 -mfma -ffp-contract=off; contraction must stay off so
 * that `p = p + x*y` is a rounded multiply followed by a rounded add).
 */
#define _GNU_SOURCE
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORC_AVX2 1
#else
#define ORC_AVX2 0
#endif

#define ORC_KC 256 /* matrixmultiply archparam::S_KC */

int orc_has_avx2_kernel(void) { return ORC_AVX2; }

/* ------------------------------------------------------------------------------------------ */
/* 1-D dots                                                                                   */
/* ------------------------------------------------------------------------------------------ */

float orc_unrolled_dot(const float *x, const float *y, size_t n)
{
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f, p5 = 0.f, p6 = 0.f, p7 = 0.f;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        p0 = p0 + x[i + 0] * y[i + 0];
        p1 = p1 + x[i + 1] * y[i + 1];
        p2 = p2 + x[i + 2] * y[i + 2];
        p3 = p3 + x[i + 3] * y[i + 3];
        p4 = p4 + x[i + 4] * y[i + 4];
        p5 = p5 + x[i + 5] * y[i + 5];
        p6 = p6 + x[i + 6] * y[i + 6];
        p7 = p7 + x[i + 7] * y[i + 7];
    }
    float sum = 0.f;
    sum = sum + (p0 + p4);
    sum = sum + (p1 + p5);
    sum = sum + (p2 + p6);
    sum = sum + (p3 + p7);
    for (; i < n; i++)
        sum = sum + x[i] * y[i];
    return sum;
}

float orc_strided_dot(const float *x, ptrdiff_t sx, const float *y, ptrdiff_t sy, size_t n)
{
    float sum = 0.f;
    for (size_t i = 0; i < n; i++)
        sum = sum + x[(ptrdiff_t)i * sx] * y[(ptrdiff_t)i * sy];
    return sum;
}

/* ndarray Ix1.dot(Ix1): unrolled form when both operands are contiguous slices. */
static float dot1(const float *x, ptrdiff_t sx, const float *y, ptrdiff_t sy, size_t n)
{
    if ((sx == 1 || n <= 1) && (sy == 1 || n <= 1))
        return orc_unrolled_dot(x, y, n);
    return orc_strided_dot(x, sx, y, sy, n);
}

/* ------------------------------------------------------------------------------------------ */
/* sgemm model                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* B packed as [k][npad] (npad multiple of 16, zero padded) — packing changes no arithmetic. */
static float *pack_b(size_t k, size_t n, const float *b, ptrdiff_t rsb, ptrdiff_t csb, size_t *npad_out)
{
    size_t npad = (n + 15) & ~(size_t)15;
    float *bp = (float *)aligned_alloc(64, (k ? k : 1) * npad * sizeof(float));
    for (size_t t = 0; t < k; t++) {
        for (size_t j = 0; j < n; j++)
            bp[t * npad + j] = b[(ptrdiff_t)t * rsb + (ptrdiff_t)j * csb];
        for (size_t j = n; j < npad; j++)
            bp[t * npad + j] = 0.f;
    }
    *npad_out = npad;
    return bp;
}

/* C[i0..i1, :] = A[i0..i1, :] * Bp with the kc-blocked sequential-FMA accumulation. */
static void gemm_rows(size_t i0, size_t i1, size_t k, size_t n, size_t npad,
                      const float *a, ptrdiff_t rsa, ptrdiff_t csa, const float *bp,
                      float *c, ptrdiff_t rsc, ptrdiff_t csc)
{
    if (k == 0) {
        for (size_t i = i0; i < i1; i++)
            for (size_t j = 0; j < n; j++)
                c[(ptrdiff_t)i * rsc + (ptrdiff_t)j * csc] = 0.f;
        return;
    }
    for (size_t kb = 0; kb < k; kb += ORC_KC) {
        size_t ke = kb + ORC_KC < k ? kb + ORC_KC : k;
        for (size_t i = i0; i < i1; i += 4) {
            size_t rows = i1 - i < 4 ? i1 - i : 4;
            for (size_t j = 0; j < npad; j += 16) {
                float acc[4][16];
#if ORC_AVX2
                if (rows == 4) { /* full 4x16 register tile: 8 independent FMA chains */
                    __m256 v00 = _mm256_setzero_ps(), v01 = v00, v10 = v00, v11 = v00;
                    __m256 v20 = v00, v21 = v00, v30 = v00, v31 = v00;
                    const float *a0 = a + (ptrdiff_t)(i + 0) * rsa, *a1 = a + (ptrdiff_t)(i + 1) * rsa;
                    const float *a2 = a + (ptrdiff_t)(i + 2) * rsa, *a3 = a + (ptrdiff_t)(i + 3) * rsa;
                    for (size_t t = kb; t < ke; t++) {
                        __m256 b0 = _mm256_load_ps(bp + t * npad + j);
                        __m256 b1 = _mm256_load_ps(bp + t * npad + j + 8);
                        __m256 av;
                        av = _mm256_broadcast_ss(a0 + (ptrdiff_t)t * csa);
                        v00 = _mm256_fmadd_ps(av, b0, v00); v01 = _mm256_fmadd_ps(av, b1, v01);
                        av = _mm256_broadcast_ss(a1 + (ptrdiff_t)t * csa);
                        v10 = _mm256_fmadd_ps(av, b0, v10); v11 = _mm256_fmadd_ps(av, b1, v11);
                        av = _mm256_broadcast_ss(a2 + (ptrdiff_t)t * csa);
                        v20 = _mm256_fmadd_ps(av, b0, v20); v21 = _mm256_fmadd_ps(av, b1, v21);
                        av = _mm256_broadcast_ss(a3 + (ptrdiff_t)t * csa);
                        v30 = _mm256_fmadd_ps(av, b0, v30); v31 = _mm256_fmadd_ps(av, b1, v31);
                    }
                    _mm256_storeu_ps(acc[0], v00); _mm256_storeu_ps(acc[0] + 8, v01);
                    _mm256_storeu_ps(acc[1], v10); _mm256_storeu_ps(acc[1] + 8, v11);
                    _mm256_storeu_ps(acc[2], v20); _mm256_storeu_ps(acc[2] + 8, v21);
                    _mm256_storeu_ps(acc[3], v30); _mm256_storeu_ps(acc[3] + 8, v31);
                } else {
                    __m256 v[4][2];
                    for (size_t r = 0; r < 4; r++) v[r][0] = v[r][1] = _mm256_setzero_ps();
                    for (size_t t = kb; t < ke; t++) {
                        __m256 b0 = _mm256_load_ps(bp + t * npad + j);
                        __m256 b1 = _mm256_load_ps(bp + t * npad + j + 8);
                        for (size_t r = 0; r < rows; r++) {
                            __m256 av = _mm256_broadcast_ss(a + (ptrdiff_t)(i + r) * rsa + (ptrdiff_t)t * csa);
                            v[r][0] = _mm256_fmadd_ps(av, b0, v[r][0]);
                            v[r][1] = _mm256_fmadd_ps(av, b1, v[r][1]);
                        }
                    }
                    for (size_t r = 0; r < rows; r++) {
                        _mm256_storeu_ps(acc[r], v[r][0]);
                        _mm256_storeu_ps(acc[r] + 8, v[r][1]);
                    }
                }
#else
                for (size_t r = 0; r < rows; r++)
                    for (size_t jj = 0; jj < 16; jj++) acc[r][jj] = 0.f;
                for (size_t t = kb; t < ke; t++)
                    for (size_t r = 0; r < rows; r++) {
                        float av = a[(ptrdiff_t)(i + r) * rsa + (ptrdiff_t)t * csa];
                        for (size_t jj = 0; jj < 16; jj++)
                            acc[r][jj] = fmaf(av, bp[t * npad + j + jj], acc[r][jj]);
                    }
#endif
                size_t cols = n > j ? (n - j < 16 ? n - j : 16) : 0;
                for (size_t r = 0; r < rows; r++) {
                    float *crow = c + (ptrdiff_t)(i + r) * rsc + (ptrdiff_t)j * csc;
                    if (kb == 0)
                        for (size_t jj = 0; jj < cols; jj++) crow[(ptrdiff_t)jj * csc] = acc[r][jj];
                    else
                        for (size_t jj = 0; jj < cols; jj++)
                            crow[(ptrdiff_t)jj * csc] = crow[(ptrdiff_t)jj * csc] + acc[r][jj];
                }
            }
        }
    }
}

void orc_sgemm(size_t m, size_t k, size_t n,
               const float *a, ptrdiff_t rsa, ptrdiff_t csa,
               const float *b, ptrdiff_t rsb, ptrdiff_t csb,
               float *c, ptrdiff_t rsc, ptrdiff_t csc)
{
    if (m == 0 || n == 0) return;
    size_t npad;
    float *bp = pack_b(k, n, b, rsb, csb, &npad);
    gemm_rows(0, m, k, n, npad, a, rsa, csa, bp, c, rsc, csc);
    free(bp);
}

/* ------------------------------------------------------------------------------------------ */
/* distances and argmin                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* Prepared codebook slice: packed transpose for the GEMM and the centroid norms. */
typedef struct {
    size_t k, dsub, kpad;
    float *bp; /* [dsub][kpad] = c^T packed */
    float *cs; /* [k] */
} prep_t;

static void prep_init(prep_t *p, const float *c, size_t k, size_t dsub)
{
    p->k = k;
    p->dsub = dsub;
    /* other.t() : B[t][j] = c[j][t] -> rsb = 1, csb = dsub  (linalg.rs:170) */
    p->bp = pack_b(dsub, k, c, 1, (ptrdiff_t)dsub, &p->kpad);
    p->cs = (float *)malloc((k ? k : 1) * sizeof(float));
    for (size_t j = 0; j < k; j++) /* linalg.rs:168 */
        p->cs[j] = orc_unrolled_dot(c + j * dsub, c + j * dsub, dsub);
}

static void prep_free(prep_t *p)
{
    free(p->bp);
    free(p->cs);
}

/* dist[i - i0, j] for rows i0..i1 of a (possibly strided) sub-matrix.  linalg.rs:157-179 */
static void sqdist_rows(const prep_t *p, const float *x, ptrdiff_t rsx, ptrdiff_t csx,
                        size_t i0, size_t i1, float *dist, float *xs_tmp)
{
    size_t k = p->k, dsub = p->dsub;
    for (size_t i = i0; i < i1; i++) { /* linalg.rs:167 */
        const float *row = x + (ptrdiff_t)i * rsx;
        xs_tmp[i - i0] = dot1(row, csx, row, csx, dsub);
    }
    gemm_rows(0, i1 - i0, dsub, k, p->kpad, x + (ptrdiff_t)i0 * rsx, rsx, csx, p->bp,
              dist, (ptrdiff_t)k, 1); /* linalg.rs:170 */
    for (size_t i = 0; i < i1 - i0; i++) { /* linalg.rs:171-176 */
        float *drow = dist + i * k;
        float xsi = xs_tmp[i];
        for (size_t j = 0; j < k; j++)
            drow[j] = xsi + p->cs[j] - (drow[j] + drow[j]);
    }
}

void orc_sqdist_batch(const float *x, size_t n, size_t ldx,
                      const float *c, size_t k, size_t dsub, float *dist)
{
    prep_t p;
    prep_init(&p, c, k, dsub);
    float *xs = (float *)malloc((n ? n : 1) * sizeof(float));
    sqdist_rows(&p, x, (ptrdiff_t)ldx, 1, 0, n, dist, xs);
    free(xs);
    prep_free(&p);
}

void orc_sqdist_vec(const float *x, const float *c, size_t k, size_t dsub, float *dist)
{
    float self_sqn = orc_unrolled_dot(x, x, dsub); /* linalg.rs:136 */
    for (size_t j = 0; j < k; j++) {
        float other_sqn = orc_unrolled_dot(c + j * dsub, c + j * dsub, dsub); /* :137 */
        float dp = orc_unrolled_dot(c + j * dsub, x, dsub);                   /* :141 mat-vec */
        dist[j] = self_sqn + other_sqn - (dp + dp);                           /* :143 */
    }
}

/* OrderedFloat "a < b": NaN is the greatest value and equal to itself (ordered-float 2). */
static inline int of_less(float a, float b)
{
    if (a < b) return 1;
    if (b != b && a == a) return 1;
    return 0;
}

/* enumerate().min_by_key(OrderedFloat): first index among equal minima (kmeans.rs:150-155). */
static inline size_t argmin_row(const float *d, size_t k)
{
    size_t best = 0;
    float bv = d[0];
    for (size_t j = 1; j < k; j++)
        if (of_less(d[j], bv)) {
            bv = d[j];
            best = j;
        }
    return best;
}

#define ORC_TEMP_ROWS ((size_t)1 << 18) /* bound on the [rows,k] temporary; results unaffected */

static void assignments_strided(const prep_t *p, const float *x, size_t n, ptrdiff_t rsx,
                                ptrdiff_t csx, uint64_t *assign, ptrdiff_t as)
{
    if (n == 0) return;
    size_t chunk = n < ORC_TEMP_ROWS ? n : ORC_TEMP_ROWS;
    float *dist = (float *)malloc(chunk * (p->k ? p->k : 1) * sizeof(float));
    float *xs = (float *)malloc(chunk * sizeof(float));
    for (size_t i0 = 0; i0 < n; i0 += chunk) {
        size_t i1 = i0 + chunk < n ? i0 + chunk : n;
        sqdist_rows(p, x, rsx, csx, i0, i1, dist, xs);
        for (size_t i = i0; i < i1; i++) /* kmeans.rs:149-156 */
            assign[(ptrdiff_t)i * as] = (uint64_t)argmin_row(dist + (i - i0) * p->k, p->k);
    }
    free(dist);
    free(xs);
}

void orc_cluster_assignments(const float *x, size_t n, size_t ldx,
                             const float *c, size_t k, size_t dsub, uint64_t *assign)
{
    prep_t p;
    prep_init(&p, c, k, dsub);
    assignments_strided(&p, x, n, (ptrdiff_t)ldx, 1, assign, 1);
    prep_free(&p);
}

uint64_t orc_cluster_assignment(const float *c, size_t k, size_t dsub, const float *x)
{
    float *dist = (float *)malloc((k ? k : 1) * sizeof(float));
    orc_sqdist_vec(x, c, k, dsub, dist);
    size_t best = argmin_row(dist, k);
    free(dist);
    return (uint64_t)best;
}

/* ------------------------------------------------------------------------------------------ */
/* k-means                                                                                    */
/* ------------------------------------------------------------------------------------------ */

void orc_update_centroids(float *c, size_t k, size_t dsub,
                          const float *x, size_t n, size_t ldx, const uint64_t *assign)
{
    memset(c, 0, k * dsub * sizeof(float)); /* kmeans.rs:181 */
    float *counts = (float *)calloc(k ? k : 1, sizeof(float));
    for (size_t i = 0; i < n; i++) { /* kmeans.rs:185-189, row order */
        float *cr = c + assign[i] * dsub;
        const float *xr = x + i * ldx;
        for (size_t t = 0; t < dsub; t++) cr[t] = cr[t] + xr[t];
        counts[assign[i]] = counts[assign[i]] + 1.0f;
    }
    for (size_t j = 0; j < k; j++) /* kmeans.rs:191-197 */
        if (counts[j] > 0.f)
            for (size_t t = 0; t < dsub; t++) c[j * dsub + t] = c[j * dsub + t] / counts[j];
    free(counts);
}

float orc_mean_squared_error(const float *c, size_t k, size_t dsub,
                             const float *x, size_t n, size_t ldx, const uint64_t *assign)
{
    (void)k;
    float sse = 0.f;
    for (size_t i = 0; i < n; i++) /* kmeans.rs:344-357, row-major iteration */
        for (size_t t = 0; t < dsub; t++) {
            float v = c[assign[i] * dsub + t] - x[i * ldx + t];
            sse = sse + v * v;
        }
    return sse / (float)(n * dsub); /* kmeans.rs:359 */
}

float orc_kmeans_iteration(float *c, size_t k, size_t dsub, const float *x, size_t n, size_t ldx)
{
    uint64_t *assign = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
    orc_cluster_assignments(x, n, ldx, c, k, dsub, assign);         /* kmeans.rs:319 */
    orc_update_centroids(c, k, dsub, x, n, ldx, assign);            /* kmeans.rs:320-325 */
    float loss = orc_mean_squared_error(c, k, dsub, x, n, ldx, assign); /* kmeans.rs:326 */
    free(assign);
    return loss;
}

float orc_kmeans_with_centroids(float *c, size_t k, size_t dsub,
                                const float *x, size_t n, size_t ldx, size_t n_iterations)
{
    float loss = 0.f;
    for (size_t iter = 0;; iter++) { /* kmeans.rs:279-284 */
        loss = orc_kmeans_iteration(c, k, dsub, x, n, ldx);
        if (iter + 1 >= n_iterations) return loss; /* NIterationsCondition, kmeans.rs:100-103 */
    }
}

/* ------------------------------------------------------------------------------------------ */
/* validation                                                                                 */
/* ------------------------------------------------------------------------------------------ */

int orc_check_quantizer_invariants(size_t n_subquantizers, uint32_t n_subquantizer_bits,
                                   size_t n_iterations, size_t n_attempts,
                                   size_t n_rows, size_t n_cols, uint64_t *detail)
{
    if (n_subquantizers == 0 || n_subquantizers > n_cols) { /* pq.rs:70-75 */
        if (detail) *detail = n_cols;
        return ORC_ERR_N_SUBQUANTIZERS_RANGE;
    }
    /* (nrows as f64).log2().trunc() as u32 — saturating cast maps -inf (0 rows) to 0. pq.rs:77 */
    uint32_t max_bits = n_rows == 0 ? 0u : (uint32_t)trunc(log2((double)n_rows));
    if (n_subquantizer_bits == 0 || n_subquantizer_bits > max_bits) { /* pq.rs:78-82 */
        if (detail) *detail = max_bits;
        return ORC_ERR_N_SUBQUANTIZER_BITS;
    }
    if (n_cols % n_subquantizers != 0) return ORC_ERR_NUMBER_SUBQUANTIZERS; /* pq.rs:84-89 */
    if (n_iterations == 0) return ORC_ERR_N_ITERATIONS;                     /* pq.rs:91-93 */
    if (n_attempts == 0) return ORC_ERR_N_ATTEMPTS;                         /* pq.rs:95-97 */
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* quantize / reconstruct                                                                     */
/* ------------------------------------------------------------------------------------------ */

static inline void store_code(void *codes, int width, ptrdiff_t idx, uint64_t v)
{
    /* `assignment.as_()` : truncating usize -> I cast (primitives.rs:100) */
    switch (width) {
    case 1: ((uint8_t *)codes)[idx] = (uint8_t)v; break;
    case 2: ((uint16_t *)codes)[idx] = (uint16_t)v; break;
    case 4: ((uint32_t *)codes)[idx] = (uint32_t)v; break;
    default: ((uint64_t *)codes)[idx] = v; break;
    }
}

static inline uint64_t load_code(const void *codes, int width, ptrdiff_t idx)
{
    switch (width) {
    case 1: return ((const uint8_t *)codes)[idx];
    case 2: return ((const uint16_t *)codes)[idx];
    case 4: return ((const uint32_t *)codes)[idx];
    default: return ((const uint64_t *)codes)[idx];
    }
}

typedef struct {
    const prep_t *preps;
    size_t M, dsub;
    const float *x;
    ptrdiff_t rsx, csx;
    void *codes;
    int code_width;
    ptrdiff_t crs, ccs;
    size_t i0, i1;
} qb_job_t;

static void *qb_worker(void *arg)
{
    qb_job_t *jb = (qb_job_t *)arg;
    size_t n = jb->i1 - jb->i0;
    if (n == 0) return NULL;
    uint64_t *assign = (uint64_t *)malloc(n * sizeof(uint64_t));
    for (size_t m = 0; m < jb->M; m++) { /* primitives.rs:90-103, sequential over subquantizers */
        const float *sub = jb->x + (ptrdiff_t)jb->i0 * jb->rsx + (ptrdiff_t)(m * jb->dsub) * jb->csx;
        assignments_strided(&jb->preps[m], sub, n, jb->rsx, jb->csx, assign, 1);
        for (size_t i = 0; i < n; i++)
            store_code(jb->codes, jb->code_width,
                       (ptrdiff_t)(jb->i0 + i) * jb->crs + (ptrdiff_t)m * jb->ccs, assign[i]);
    }
    free(assign);
    return NULL;
}

typedef struct {
    size_t i0, i1, k, n, npad;
    const float *a;
    ptrdiff_t rsa, csa;
    const float *bp;
    float *c;
    ptrdiff_t rsc, csc;
} gemm_job_t;

static void *gemm_worker(void *arg)
{
    gemm_job_t *g = (gemm_job_t *)arg;
    gemm_rows(g->i0, g->i1, g->k, g->n, g->npad, g->a, g->rsa, g->csa, g->bp, g->c, g->rsc, g->csc);
    return NULL;
}

/* Row-sharded sgemm (identical per-element arithmetic). */
static void sgemm_threads(size_t m, size_t k, size_t n,
                          const float *a, ptrdiff_t rsa, ptrdiff_t csa,
                          const float *b, ptrdiff_t rsb, ptrdiff_t csb,
                          float *c, ptrdiff_t rsc, ptrdiff_t csc, int n_threads)
{
    if (m == 0 || n == 0) return;
    if (n_threads < 1) n_threads = 1;
    size_t npad;
    float *bp = pack_b(k, n, b, rsb, csb, &npad);
    pthread_t th[256];
    gemm_job_t jobs[256];
    if (n_threads > 256) n_threads = 256;
    size_t per = ((m + (size_t)n_threads - 1) / (size_t)n_threads + 3) & ~(size_t)3;
    int used = 0;
    for (int t = 0; t < n_threads; t++) {
        size_t i0 = (size_t)t * per, i1 = i0 + per < m ? i0 + per : m;
        if (i0 >= m) break;
        jobs[t] = (gemm_job_t){i0, i1, k, n, npad, a, rsa, csa, bp, c, rsc, csc};
        used++;
    }
    for (int t = 1; t < used; t++) pthread_create(&th[t], NULL, gemm_worker, &jobs[t]);
    gemm_worker(&jobs[0]);
    for (int t = 1; t < used; t++) pthread_join(th[t], NULL);
    free(bp);
}

void orc_quantize_batch(const float *quantizers, size_t M, size_t k, size_t dsub,
                        const float *projection,
                        const float *x, size_t n, ptrdiff_t rsx, ptrdiff_t csx,
                        void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs,
                        int n_threads)
{
    size_t d = M * dsub;
    if (n == 0) return;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    float *rx = NULL;
    if (projection) { /* pq.rs:276: rx = x.dot(projection) */
        rx = (float *)malloc(n * d * sizeof(float));
        sgemm_threads(n, d, d, x, rsx, csx, projection, (ptrdiff_t)d, 1, rx, (ptrdiff_t)d, 1, n_threads);
        x = rx;
        rsx = (ptrdiff_t)d;
        csx = 1;
    }
    prep_t *preps = (prep_t *)malloc(M * sizeof(prep_t));
    for (size_t m = 0; m < M; m++) prep_init(&preps[m], quantizers + m * k * dsub, k, dsub);

    pthread_t th[256];
    qb_job_t jobs[256];
    size_t per = (n + (size_t)n_threads - 1) / (size_t)n_threads;
    int used = 0;
    for (int t = 0; t < n_threads; t++) {
        size_t i0 = (size_t)t * per, i1 = i0 + per < n ? i0 + per : n;
        if (i0 >= n) break;
        jobs[t] = (qb_job_t){preps, M, dsub, x, rsx, csx, codes, code_width, crs, ccs, i0, i1};
        used++;
    }
    for (int t = 1; t < used; t++) pthread_create(&th[t], NULL, qb_worker, &jobs[t]);
    qb_worker(&jobs[0]);
    for (int t = 1; t < used; t++) pthread_join(th[t], NULL);

    for (size_t m = 0; m < M; m++) prep_free(&preps[m]);
    free(preps);
    free(rx);
}

int orc_quantize_vector(const float *quantizers, size_t M, size_t k, size_t dsub,
                        const float *projection, const float *x, ptrdiff_t sx,
                        void *codes, int code_width)
{
    size_t d = M * dsub;
    /* primitives.rs:31-34: quantizers.len_of(Axis(1)) - 1 <= I::max_value() */
    if (code_width < 8) {
        uint64_t maxv = code_width == 1 ? 0xffull : code_width == 2 ? 0xffffull : 0xffffffffull;
        if ((uint64_t)(k - 1) > maxv) return -1;
    }
    float *v = (float *)malloc((d ? d : 1) * sizeof(float));
    if (projection) {
        /* pq.rs:293: x.dot(projection) = projection.t().dot(x): one row.dot(x) per row of R^T,
         * i.e. per (non-contiguous) column of R -> plain sequential dot unless d == 1. */
        for (size_t j = 0; j < d; j++)
            v[j] = dot1(projection + j, (ptrdiff_t)d, x, sx, d);
    } else {
        for (size_t j = 0; j < d; j++) v[j] = x[(ptrdiff_t)j * sx];
        /* a strided input view is not a slice: its norms take the sequential dot. */
    }
    int strided_in = (!projection && sx != 1 && d > 1);
    float *dist = (float *)malloc((k ? k : 1) * sizeof(float));
    for (size_t m = 0; m < M; m++) { /* primitives.rs:39-46 */
        const float *c = quantizers + m * k * dsub;
        const float *sub = v + m * dsub;
        if (!strided_in || dsub <= 1) {
            orc_sqdist_vec(sub, c, k, dsub, dist);
        } else {
            float self_sqn = orc_strided_dot(sub, 1, sub, 1, dsub);
            for (size_t j = 0; j < k; j++) {
                float other_sqn = orc_unrolled_dot(c + j * dsub, c + j * dsub, dsub);
                float dp = orc_strided_dot(c + j * dsub, 1, sub, 1, dsub);
                dist[j] = self_sqn + other_sqn - (dp + dp);
            }
        }
        store_code(codes, code_width, (ptrdiff_t)m, (uint64_t)argmin_row(dist, k));
    }
    free(dist);
    free(v);
    return 0;
}

int orc_reconstruct_batch(const float *quantizers, size_t M, size_t k, size_t dsub,
                          const float *projection,
                          const void *codes, int code_width, size_t n, ptrdiff_t crs, ptrdiff_t ccs,
                          float *out, ptrdiff_t rso, ptrdiff_t cso, int n_threads)
{
    size_t d = M * dsub;
    for (size_t i = 0; i < n; i++) /* primitives.rs:169-172 */
        for (size_t m = 0; m < M; m++) { /* primitives.rs:141-147 */
            uint64_t code = load_code(codes, code_width, (ptrdiff_t)i * crs + (ptrdiff_t)m * ccs);
            if (code >= k) return -1;
            const float *src = quantizers + (m * k + code) * dsub;
            for (size_t t = 0; t < dsub; t++)
                out[(ptrdiff_t)i * rso + (ptrdiff_t)(m * dsub + t) * cso] = src[t];
        }
    if (projection && n > 0) { /* pq.rs:323-326: out = out.dot(R.t()) then assign */
        float *tmp = (float *)malloc(n * d * sizeof(float));
        sgemm_threads(n, d, d, out, rso, cso, projection, 1, (ptrdiff_t)d, tmp, (ptrdiff_t)d, 1,
                      n_threads);
        for (size_t i = 0; i < n; i++)
            for (size_t j = 0; j < d; j++)
                out[(ptrdiff_t)i * rso + (ptrdiff_t)j * cso] = tmp[i * d + j];
        free(tmp);
    }
    return 0;
}

int orc_reconstruct(const float *quantizers, size_t M, size_t k, size_t dsub,
                    const float *projection, const void *codes, int code_width, float *out)
{
    size_t d = M * dsub;
    for (size_t m = 0; m < M; m++) {
        uint64_t code = load_code(codes, code_width, (ptrdiff_t)m);
        if (code >= k) return -1;
        memcpy(out + m * dsub, quantizers + (m * k + code) * dsub, dsub * sizeof(float));
    }
    if (projection) {
        /* pq.rs:340: reconstruction.dot(&projection.t()) = projection.dot(reconstruction):
         * one contiguous row.dot(v) per row of R -> unrolled_dot. */
        float *tmp = (float *)malloc((d ? d : 1) * sizeof(float));
        for (size_t i = 0; i < d; i++) tmp[i] = orc_unrolled_dot(projection + i * d, out, d);
        memcpy(out, tmp, d * sizeof(float));
        free(tmp);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* training                                                                                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    const float *x;
    size_t n, d, M, k, dsub, n_iterations, n_attempts;
    const float *initial;
    float *out_q, *out_loss;
    int tid, n_threads;
} train_job_t;

static void *train_worker(void *arg)
{
    train_job_t *tj = (train_job_t *)arg;
    size_t kd = tj->k * tj->dsub;
    float *cand = (float *)malloc(kd * sizeof(float));
    for (size_t m = (size_t)tj->tid; m < tj->M; m += (size_t)tj->n_threads) {
        const float *sub = tj->x + m * tj->dsub; /* pq.rs:166 */
        float best_loss = 0.f;
        for (size_t a = 0; a < tj->n_attempts; a++) { /* pq.rs:168-183 */
            memcpy(cand, tj->initial + (a * tj->M + m) * kd, kd * sizeof(float));
            float loss = orc_kmeans_with_centroids(cand, tj->k, tj->dsub, sub, tj->n, tj->d,
                                                   tj->n_iterations);
            if (a == 0 || of_less(loss, best_loss)) { /* min_by_key, first minimum. pq.rs:184-185 */
                best_loss = loss;
                memcpy(tj->out_q + m * kd, cand, kd * sizeof(float));
            }
        }
        tj->out_loss[m] = best_loss;
    }
    free(cand);
    return NULL;
}

int orc_train_pq(const float *x, size_t n, size_t d, size_t M, uint32_t n_bits,
                 size_t n_iterations, size_t n_attempts, const float *initial,
                 float *out_quantizers, float *out_loss, int n_threads)
{
    int rc = orc_check_quantizer_invariants(M, n_bits, n_iterations, n_attempts, n, d, NULL);
    if (rc != ORC_OK) return rc;
    size_t k = (size_t)1 << n_bits; /* pq.rs:233 */
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if ((size_t)n_threads > M) n_threads = (int)M;
    pthread_t th[256];
    train_job_t jobs[256];
    for (int t = 0; t < n_threads; t++)
        jobs[t] = (train_job_t){x, n, d, M, k, d / M, n_iterations, n_attempts, initial,
                                out_quantizers, out_loss, t, n_threads};
    for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, train_worker, &jobs[t]);
    train_worker(&jobs[0]);
    for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
    return ORC_OK;
}
