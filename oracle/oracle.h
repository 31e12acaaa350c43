/*
 * oracle.h — CPU restatement of reductive 0.9.0's product-quantization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under reductive_b200/ may include, link or call this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * Arithmetic model ("parity pinned on the reference's golden vectors only; large-shape bit
 * behaviour is a restatement of un-vendored third-party code" — see oracle.c header):
 *   reductive 0.9.0 + ndarray 0.15 (no `blas` feature) + matrixmultiply 0.3.x on an
 *   FMA-capable x86-64 host, kc = 256.
 */
#ifndef REDUCTIVE_ORACLE_H
#define REDUCTIVE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes mirror ReductiveError (reference src/error.rs:6-41). */
enum {
    ORC_OK = 0,
    ORC_ERR_N_ATTEMPTS = 1,               /* IncorrectNAttempts            error.rs:9  */
    ORC_ERR_N_ITERATIONS = 2,             /* IncorrectNIterations          error.rs:12 */
    ORC_ERR_N_SUBQUANTIZER_BITS = 3,      /* IncorrectNSubquantizerBits    error.rs:18 */
    ORC_ERR_NUMBER_SUBQUANTIZERS = 4,     /* IncorrectNumberSubquantizers  error.rs:25 */
    ORC_ERR_N_SUBQUANTIZERS_RANGE = 5     /* NSubquantizersOutsideRange    error.rs:35 */
};

/* ndarray 0.15 numeric_util::unrolled_dot (called from linalg.rs:110-112,136-137,167-168). */
float orc_unrolled_dot(const float *x, const float *y, size_t n);
/* ndarray 0.15 non-contiguous 1-D dot: plain sequential sum = sum + a*b. */
float orc_strided_dot(const float *x, ptrdiff_t sx, const float *y, ptrdiff_t sy, size_t n);

/* matrixmultiply 0.3 sgemm model: C[m,n] = A[m,k] * B[k,n], arbitrary element strides, beta = 0.
 * Each C element is a sequential fused-multiply-add chain over k inside blocks of kc = 256;
 * blocks are combined with a plain add.  (call sites: linalg.rs:170, pq.rs:276, pq.rs:324) */
void orc_sgemm(size_t m, size_t k, size_t n,
               const float *a, ptrdiff_t rsa, ptrdiff_t csa,
               const float *b, ptrdiff_t rsb, ptrdiff_t csb,
               float *c, ptrdiff_t rsc, ptrdiff_t csc);

/* linalg.rs:150-180: dist[i,j] = (xs[i] + cs[j]) - (dp + dp);  x rows have stride ldx (elements). */
void orc_sqdist_batch(const float *x, size_t n, size_t ldx,
                      const float *c, size_t k, size_t dsub, float *dist);
/* linalg.rs:118-148: vector vs matrix (mat-vec through unrolled_dot). */
void orc_sqdist_vec(const float *x, const float *c, size_t k, size_t dsub, float *dist);

/* kmeans.rs:133-159 (Axis(0)).  assign[i] = first index of the minimum; NaN ranks largest. */
void orc_cluster_assignments(const float *x, size_t n, size_t ldx,
                             const float *c, size_t k, size_t dsub, uint64_t *assign);
/* kmeans.rs:111-126 */
uint64_t orc_cluster_assignment(const float *c, size_t k, size_t dsub, const float *x);
/* kmeans.rs:166-198 (sequential scatter-add, float counts, empty clusters left at zero). */
void orc_update_centroids(float *c, size_t k, size_t dsub,
                          const float *x, size_t n, size_t ldx, const uint64_t *assign);
/* kmeans.rs:330-360 */
float orc_mean_squared_error(const float *c, size_t k, size_t dsub,
                             const float *x, size_t n, size_t ldx, const uint64_t *assign);
/* kmeans.rs:308-327: assign -> update -> mse(new centroids, old assignments). */
float orc_kmeans_iteration(float *c, size_t k, size_t dsub, const float *x, size_t n, size_t ldx);
/* kmeans.rs:263-288 with NIterationsCondition (kmeans.rs:97-104). */
float orc_kmeans_with_centroids(float *c, size_t k, size_t dsub,
                                const float *x, size_t n, size_t ldx, size_t n_iterations);

/* pq.rs:63-100.  Returns ORC_OK or one of ORC_ERR_*; *detail receives max_subquantizer_bits /
 * max_subquantizers where the reference's error carries one. */
int orc_check_quantizer_invariants(size_t n_subquantizers, uint32_t n_subquantizer_bits,
                                   size_t n_iterations, size_t n_attempts,
                                   size_t n_rows, size_t n_cols, uint64_t *detail);

/* pq.rs:256-283 -> primitives.rs:64-104.  quantizers [M,k,dsub] contiguous; projection [d,d]
 * row-major or NULL; x [n,d] with element strides (rsx, csx); codes written as code_width-byte
 * little-endian integers (truncating cast, primitives.rs:100) at [i*crs + m*ccs] (element strides).
 * scratch-free: allocates internally.  n_threads > 1 shards rows (results are row-independent). */
void orc_quantize_batch(const float *quantizers, size_t M, size_t k, size_t dsub,
                        const float *projection,
                        const float *x, size_t n, ptrdiff_t rsx, ptrdiff_t csx,
                        void *codes, int code_width, ptrdiff_t crs, ptrdiff_t ccs,
                        int n_threads);
/* pq.rs:285-298 -> primitives.rs:14-49.  Returns 0, or -1 when k-1 does not fit code_width
 * (the reference panics, primitives.rs:31-34). */
int orc_quantize_vector(const float *quantizers, size_t M, size_t k, size_t dsub,
                        const float *projection, const float *x, ptrdiff_t sx,
                        void *codes, int code_width);
/* traits.rs:109-117 -> pq.rs:309-327 -> primitives.rs:150-173 (+ recon . R^T).  Returns 0, or -1 on
 * an out-of-range code (the reference panics on the ndarray index). */
int orc_reconstruct_batch(const float *quantizers, size_t M, size_t k, size_t dsub,
                          const float *projection,
                          const void *codes, int code_width, size_t n, ptrdiff_t crs, ptrdiff_t ccs,
                          float *out, ptrdiff_t rso, ptrdiff_t cso, int n_threads);
/* pq.rs:329-343 -> primitives.rs:110-148 (+ v . R^T through the non-contiguous dot). */
int orc_reconstruct(const float *quantizers, size_t M, size_t k, size_t dsub,
                    const float *projection, const void *codes, int code_width, float *out);

/* pq.rs:201-249 / :144-188 with the initial centroids supplied by the caller
 * (initial [n_attempts, M, k, dsub]); one thread per subquantizer up to n_threads (Rayon at
 * pq.rs:226).  x is [n,d] contiguous.  out_quantizers [M,k,dsub], out_loss [M]. */
int orc_train_pq(const float *x, size_t n, size_t d, size_t M, uint32_t n_bits,
                 size_t n_iterations, size_t n_attempts, const float *initial,
                 float *out_quantizers, float *out_loss, int n_threads);

/* Build-time facts, for the bench JSON. */
int orc_has_avx2_kernel(void);

#ifdef __cplusplus
}
#endif
#endif
