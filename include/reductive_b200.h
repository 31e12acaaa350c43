/*
 * reductive_b200.h — C ABI of the B200-native product-quantization hot path.
 *
 * This is the drop-in boundary: every entry point is what reductive's FFI for this path would bind.
 * Each declaration cites the reference interface it replaces (paths relative to /root/reference).
 * Plain pointers and sizes only; no torch / ndarray types cross this boundary.
 *
 * Conventions
 *   - Strides are in ELEMENTS (like ndarray), may be any non-zero value for inputs; views with
 *     col_stride != 1 are packed on the device before the kernels run.
 *   - `mem_kind` says where the data pointers of that call live (RB_MEM_HOST: ordinary (pageable) or pinned host
 *     memory, copied in chunks overlapped with the kernels -- pageable memory goes through pinned staging buffers
 *     the library owns, so it reaches about the pinned transfer rate; RB_MEM_DEVICE: device memory of the
 *     CURRENT CUDA device, no copies).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).  Device-memory calls
 *     are asynchronous on that stream; host-memory calls return after the result is in host memory.
 *   - A handle belongs to the device that was current when it was created (rb_pq_create) or to the device list it
 *     was given (rb_pq_create_multi: host-memory batches are split over the devices from this one process).  Handles are immutable after creation: concurrent calls on one
 *     handle from several threads / streams are legal (Pq is Sync in the reference, pq.rs:28).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point returns
 *     RB_ERR_NO_DEVICE / RB_ERR_CUDA.
 *   - Codes are little-endian unsigned integers of `code_width` bytes (1, 2, 4 or 8) — the reference's
 *     generic index type I (traits.rs:77-80).
 */
#ifndef REDUCTIVE_B200_H
#define REDUCTIVE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: + rb_qstore_*, rb_set_gram_algo, rb_kmeans_dist_peer_window (additions only; version-2 callers keep working) */
#define RB_ABI_VERSION 3

/* Status codes.  1..6 mirror ReductiveError (src/error.rs:6-41); 16.. are the reference's panics
 * (assert sites: primitives.rs:25-34,74-87,123-135,159-167; pq.rs:39-55; kmeans.rs:61-71,269-277),
 * which a Rust shim re-asserts before calling so that panic behaviour is preserved. */
typedef enum rb_status {
    RB_OK = 0,
    RB_ERR_N_ATTEMPTS = 1,            /* ReductiveError::IncorrectNAttempts            error.rs:9  */
    RB_ERR_N_ITERATIONS = 2,          /* ReductiveError::IncorrectNIterations          error.rs:12 */
    RB_ERR_N_SUBQUANTIZER_BITS = 3,   /* ReductiveError::IncorrectNSubquantizerBits    error.rs:18 */
    RB_ERR_NUMBER_SUBQUANTIZERS = 4,  /* ReductiveError::IncorrectNumberSubquantizers  error.rs:25 */
    RB_ERR_N_SUBQUANTIZERS_RANGE = 5, /* ReductiveError::NSubquantizersOutsideRange    error.rs:35 */
    RB_ERR_CONSTRUCT_RNG = 6,         /* ReductiveError::ConstructRng                  error.rs:40 */
    RB_ERR_SHAPE = 16,                /* shape / length mismatch (reference: assert! panic)        */
    RB_ERR_CODE_TYPE = 17,            /* k-1 does not fit the code type (primitives.rs:31-34)      */
    RB_ERR_CODE_RANGE = 18,           /* code >= k in reconstruct (reference: ndarray index panic) */
    RB_ERR_INVALID = 19,              /* NULL pointer / bad enum                                   */
    RB_ERR_K_MEANS_K = 20,            /* k == 0 or k >= n instances (kmeans.rs:61-67)              */
    RB_ERR_CUDA = 32,                 /* a CUDA runtime call or kernel failed                      */
    RB_ERR_NO_DEVICE = 33,            /* no CUDA device visible                                    */
    RB_ERR_UNSUPPORTED = 34,          /* shape outside what the kernels cover (message says what)  */
    RB_ERR_NCCL = 35                  /* an NCCL call failed / libnccl.so.2 could not be loaded    */
} rb_status;

typedef enum rb_mem_kind { RB_MEM_HOST = 0, RB_MEM_DEVICE = 1 } rb_mem_kind;

/* Which encode kernel quantize_batch / k-means assignment uses.  Both give bit-identical codes. */
typedef enum rb_encode_algo {
    RB_ENCODE_AUTO = 0,   /* tensor path when the shape allows it, otherwise exact SIMT */
    RB_ENCODE_EXACT = 1,  /* FP32 SIMT kernel evaluating the reference's FMA chain directly */
    RB_ENCODE_TENSOR = 2  /* tcgen05 candidate pass + exact recheck of near-ties (64 < k <= 256) */
} rb_encode_algo;

/* Which kernel applies the OPQ rotation (x.R before the argmin, .R^T after the gather; pq.rs:276,324) in the
 * batch entry points.  Codes are bit-identical under all three; rotated reconstructions are bit-identical to the
 * reference's FP32 GEMM under EXACT and within 1e-5 of it under TENSOR. */
typedef enum rb_project_algo {
    RB_PROJECT_AUTO = 0,   /* tensor path for large aligned batches, otherwise exact */
    RB_PROJECT_EXACT = 1,  /* FP32 SIMT GEMM in the reference's summation order */
    RB_PROJECT_TENSOR = 2  /* tcgen05 two-limb FP16 GEMM (d % 4 == 0, 32 <= d <= 4096, n >= 1024) */
} rb_project_algo;

/* Opaque product quantizer: reference `Pq<f32>` (src/pq/pq.rs:29-32) resident on one GPU. */
typedef struct rb_pq rb_pq;

/* Thread-local description of the last non-OK status returned on this thread. */
const char *rb_last_error_message(void);
int rb_abi_version(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t rb_kernel_launch_count(void);
/* Process-wide default for RB_ENCODE_AUTO resolution (tests force each path). */
rb_status rb_set_encode_algo(int algo);
/* Scratch buffers (workspaces, flagged-pair lists, sort buffers) come from a per-device memory pool the library
 * keeps across calls; this hands the unused part back to the driver (current device). */
rb_status rb_release_scratch(void);
/* Process-wide choice of the rotation kernel (rb_project_algo). */
rb_status rb_set_project_algo(int algo);

/* Centroid-update summation order of the k-means entry points (process-wide).  ordered != 0 (default 1):
 * rows of a cluster are added sequentially in row order exactly like kmeans.rs:185-189, so sums are
 * bit-identical to the reference's (1, 2: stable sort + chain kernels; 3: the training loops stream
 * subquantizer-major slabs instead -- the same bits, measured slower, kept as an experiment); 0: shared-memory
 * atomics, order unspecified. */
rb_status rb_set_kmeans_update(int ordered);

/* ---- Pq construction and accessors ----------------------------------------------------------- */

/* Pq::new(projection, quantizers)  pq.rs:38-61.  quantizers: HOST [M,k,dsub] contiguous;
 * projection: HOST [d,d] row-major (d = M*dsub) or NULL.  RB_ERR_SHAPE when any extent is 0. */
rb_status rb_pq_create(const float *quantizers, size_t n_subquantizers, size_t n_centroids,
                       size_t subquantizer_dim, const float *projection_or_null, rb_pq **out);
/* The same quantizer replicated on several devices of this process.  The handle lives on devices[0];
 * rb_pq_quantize_batch / rb_pq_reconstruct_batch with RB_MEM_HOST then split the rows into contiguous blocks over
 * the devices (one host thread and one copy / compute pipeline per device, no collective: rows are independent).
 * Device-memory calls use the replica of devices[0]. */
rb_status rb_pq_create_multi(const float *quantizers, size_t n_subquantizers, size_t n_centroids,
                             size_t subquantizer_dim, const float *projection_or_null, const int *devices,
                             int n_devices, rb_pq **out);
int rb_pq_n_devices(const rb_pq *pq);
/* Host threads per pageable <-> pinned staging copy of the host-memory pipelines (default 4; process-wide). */
rb_status rb_set_host_copy_threads(int n_threads);
void rb_pq_destroy(rb_pq *pq);
size_t rb_pq_quantized_len(const rb_pq *pq);         /* QuantizeVector::quantized_len     pq.rs:300 */
size_t rb_pq_reconstructed_len(const rb_pq *pq);     /* Reconstruct::reconstructed_len    pq.rs:345 */
size_t rb_pq_n_quantizer_centroids(const rb_pq *pq); /* Pq::n_quantizer_centroids         pq.rs:103 */
int rb_pq_has_projection(const rb_pq *pq);           /* Pq::projection().is_some()        pq.rs:108 */
rb_status rb_pq_subquantizers(const rb_pq *pq, float *out_host);  /* Pq::subquantizers pq.rs:191 */
rb_status rb_pq_projection(const rb_pq *pq, float *out_host);     /* Pq::projection    pq.rs:108 */

/* ---- QuantizeVector (src/pq/traits.rs:75-99; impl pq.rs:252-303) ------------------------------ */

/* quantize_batch_into  pq.rs:268-283 -> primitives.rs:64-104.  x [n, d] f32, codes [n, M].
 * No check that k-1 fits code_width (the reference's batch path truncates, primitives.rs:100).
 * A quantizer with a projection rotates the rows first (pq.rs:276) with the kernel rb_set_project_algo selects;
 * the codes are the same under every choice. */
rb_status rb_pq_quantize_batch(const rb_pq *pq, const float *x, size_t n, ptrdiff_t x_row_stride,
                               ptrdiff_t x_col_stride, void *codes, int code_width,
                               ptrdiff_t code_row_stride, ptrdiff_t code_col_stride,
                               int mem_kind, void *stream);
/* quantize_vector  pq.rs:285-298 -> primitives.rs:14-49 (mat-vec arithmetic, differs from the batch
 * path in rounding).  RB_ERR_CODE_TYPE when k-1 > max of the code type. */
rb_status rb_pq_quantize_vector(const rb_pq *pq, const float *x, ptrdiff_t x_stride, void *codes,
                                int code_width, ptrdiff_t code_stride, int mem_kind, void *stream);

/* ---- Reconstruct (src/pq/traits.rs:102-156; impl pq.rs:305-348) ------------------------------- */

/* reconstruct_batch_into  pq.rs:309-327 -> primitives.rs:150-173 (+ out = out . R^T).  RB_ERR_CODE_RANGE
 * if any code >= k (detected on the device; output rows with a bad code are unspecified).  Reporting that
 * status needs the kernel's result, so a device-memory call synchronises `stream` -- except when no code value can
 * be out of range (code_width 1, k = 256), where it stays asynchronous. */
rb_status rb_pq_reconstruct_batch(const rb_pq *pq, const void *codes, int code_width, size_t n,
                                  ptrdiff_t code_row_stride, ptrdiff_t code_col_stride, float *out,
                                  ptrdiff_t out_row_stride, ptrdiff_t out_col_stride, int mem_kind,
                                  void *stream);
/* reconstruct_into  pq.rs:329-343 -> primitives.rs:110-148. */
rb_status rb_pq_reconstruct(const rb_pq *pq, const void *codes, int code_width, ptrdiff_t code_stride,
                            float *out, ptrdiff_t out_stride, int mem_kind, void *stream);

/* ---- TrainPq (src/pq/traits.rs:15-72; impl pq.rs:196-250) and k-means (src/kmeans.rs) --------- */

/* Pq::check_quantizer_invariants  pq.rs:63-100 (host only).  *detail (may be NULL) receives
 * max_subquantizer_bits / max_subquantizers for the two errors that carry one. */
rb_status rb_check_quantizer_invariants(size_t n_subquantizers, uint32_t n_subquantizer_bits,
                                        size_t n_iterations, size_t n_attempts, size_t n_rows,
                                        size_t n_cols, uint64_t *detail);

/* Length (in floats) of the packed per-iteration accumulator:
 * sums [M,k,dsub] | counts [M,k] | sum of squared norms [M].  This is the all-reduce payload of
 * data-parallel k-means. */
size_t rb_kmeans_packed_len(size_t n_subquantizers, size_t n_centroids, size_t subquantizer_dim);

/* First half of KMeansIteration::kmeans_iteration for all M subquantizers at once (kmeans.rs:319-325):
 * cluster_assignments (kmeans.rs:133-159) of the LOCAL rows against `centroids`, then the scatter-add
 * of update_centroids (kmeans.rs:181-189) into `packed` (zeroed by this call).  All pointers are
 * DEVICE memory.  x: [n_local, d] with row stride x_row_stride (col stride 1). */
rb_status rb_kmeans_assign_accumulate(const float *x, size_t n_local, ptrdiff_t x_row_stride,
                                      const float *centroids, size_t n_subquantizers,
                                      size_t n_centroids, size_t subquantizer_dim, float *packed,
                                      void *stream);
/* The two halves of the call above, for callers that overlap or chain them (data-parallel training):
 *   rb_kmeans_assign      cluster_assignments (kmeans.rs:133-159) of the local rows for all M subquantizers into
 *                         `codes`: DEVICE, column-major [M][rb_kmeans_code_pitch(n_local)] elements of
 *                         rb_kmeans_code_width(k) bytes (1 for k <= 256, else 4);
 *   rb_kmeans_accumulate  the scatter-add of update_centroids (kmeans.rs:181-189) into `packed`, CONTINUING from
 *                         `packed_before` (DEVICE, rb_kmeans_packed_len floats, or NULL = from zero): the sums of the
 *                         rows that precede these n_local rows in the reference's row order.  kmeans.rs:185-189 adds
 *                         the rows of a cluster sequentially in f32, so passing the running sums from rank to rank
 *                         (rank r's rows follow rank r-1's) keeps data-parallel training bit-identical to one
 *                         sequential pass.  A non-NULL packed_before needs the ordered update (default) with
 *                         k <= 256 and subquantizer_dim <= 32; RB_ERR_UNSUPPORTED otherwise. */
size_t rb_kmeans_code_pitch(size_t n_local);
int rb_kmeans_code_width(size_t n_centroids);
rb_status rb_kmeans_assign(const float *x, size_t n_local, ptrdiff_t x_row_stride, const float *centroids,
                           size_t n_subquantizers, size_t n_centroids, size_t subquantizer_dim, void *codes,
                           void *stream);
rb_status rb_kmeans_accumulate(const float *x, size_t n_local, ptrdiff_t x_row_stride, const void *codes,
                               size_t n_subquantizers, size_t n_centroids, size_t subquantizer_dim,
                               const float *packed_before, float *packed, void *stream);
/* rb_kmeans_assign followed by rb_kmeans_accumulate with an internal assignment buffer. */
rb_status rb_kmeans_assign_accumulate_from(const float *x, size_t n_local, ptrdiff_t x_row_stride,
                                           const float *centroids, size_t n_subquantizers, size_t n_centroids,
                                           size_t subquantizer_dim, const float *packed_before, float *packed,
                                           void *stream);
/* Second half (kmeans.rs:191-197 + mean_squared_error kmeans.rs:330-360): divide non-empty clusters,
 * leave empty clusters at zero, write the new centroids and, if loss_or_null != NULL, the per-
 * subquantizer mean squared error of the new centroids under the old assignments (computed from the
 * accumulated moments in FP64).  `packed` holds the (all-reduced) sums over n_total rows.  DEVICE memory. */
rb_status rb_kmeans_finalize(const float *packed, size_t n_subquantizers, size_t n_centroids,
                             size_t subquantizer_dim, uint64_t n_total, float *centroids,
                             float *loss_or_null, void *stream);

/* TrainPq::train_pq_using for Pq  pq.rs:201-249 with the initial centroids supplied by the caller
 * (initial: HOST [n_attempts, M, k, dsub]; the reference draws them with RandomInstanceCentroids,
 * kmeans.rs:52-87, whose HashSet iteration order is not reproducible, so they stay a host-side input).
 * instances: [n, d] (mem_kind).  loss_out: HOST [M] or NULL.  Best of n_attempts by loss (pq.rs:183-187). */
rb_status rb_pq_train(const float *instances, size_t n, size_t d, ptrdiff_t row_stride,
                      ptrdiff_t col_stride, size_t n_subquantizers, uint32_t n_subquantizer_bits,
                      size_t n_iterations, size_t n_attempts, const float *initial_centroids,
                      float *loss_out, int mem_kind, void *stream, rb_pq **out);

/* ---- projection (pq.rs:276, pq.rs:323-326; opq.rs:68,173) ------------------------------------- */

/* out[n,d] = x[n,d] . R (transpose_r = 0) or x . R^T (transpose_r = 1) with the reference's FP32
 * accumulation order (sequential FMA over k, kc = 256 blocks).  DEVICE memory; out row-major. */
rb_status rb_project_rows(const float *x, size_t n, size_t d, ptrdiff_t x_row_stride,
                          ptrdiff_t x_col_stride, const float *r_dev, int transpose_r, float *out,
                          void *stream);

/* ---- Opq / GaussianOpq training (opq.rs:46-209, gaussian_opq.rs:33-68) ------------------------------ */

/* Gram matrices of Opq training (rb_covariance, the X^T.Y^ of rb_opq_train_iteration): 0 = auto (tensor cores for
 * n >= 4096 rows and 16 <= d <= 512, FP32 CUDA cores otherwise), 1 = FP32 CUDA cores, 2 = tensor cores (UNSUPPORTED
 * outside that range).  Both are within 1e-5 of the exact product relative to its largest element. */
rb_status rb_set_gram_algo(int algo);
/* Covariance::covariance over observation axis 0 (linalg.rs:23-44) of x [n, d] (DEVICE, row stride x_row_stride):
 * cov_out DEVICE [d, d].  RB_ERR_SHAPE for n == 0 (the reference asserts).  The d x d eigendecomposition that
 * follows in Opq::create_projection_matrix (opq.rs:120-135) stays on host LAPACK, as in the reference. */
rb_status rb_covariance(const float *x, size_t n, size_t d, ptrdiff_t x_row_stride, float *cov_out, void *stream);

/* The device part of Opq::train_iteration (opq.rs:161-189) for all M subquantizers:
 *   rx = x . projection (opq.rs:173, reference summation order); one kmeans_iteration per subquantizer on rx
 *   (opq.rs:174,191-209; centroids updated in place); quantize -> reconstruct round trip with the new centroids
 *   (opq.rs:180-182); xty_out = x^T . reconstructed (opq.rs:187).
 * The caller takes the SVD of xty_out on host LAPACK and passes U.V^T as the next projection (opq.rs:187-188).
 * All pointers DEVICE: x [n, d] (row stride), projection [d, d], centroids [M, k, dsub], xty_out [d, d]. */
rb_status rb_opq_train_iteration(const float *x, size_t n, size_t d, ptrdiff_t x_row_stride, const float *projection,
                                 float *centroids, size_t n_subquantizers, size_t n_centroids, float *xty_out,
                                 void *stream);

/* ---- caller-side quantized storage ----------------------------------------------------
 * finalfusion keeps a trained Pq<f32> as a "quantized array": the quantizer, the [n, M] u8 codes produced by
 * QuantizeVector::quantize_batch (traits.rs:77-87) and optionally one norm per row (rows are l2-normalised before
 * quantisation); its lookup is  embedding(i) = Reconstruct::reconstruct(codes[i]) * norm[i]  (traits.rs:102-156,
 * pq.rs:303-347).  That type is outside /root/reference; these entry points are what a binding of it would call.
 * The store keeps codes and norms resident in HBM and BORROWS the quantizer (destroy the store first).
 * u8 codes only: a quantizer with more than 256 centroids answers RB_ERR_CODE_TYPE. */
typedef struct rb_qstore rb_qstore;
/* codes: [n, M] u8 with row stride code_row_stride (elements, unit column stride); norms_or_null: [n] f32; both in
 * mem_kind memory.  Copies them; a code >= n_centroids answers RB_ERR_CODE_RANGE.  Synchronises `stream`. */
rb_status rb_qstore_create(const rb_pq *pq, const uint8_t *codes, size_t n, ptrdiff_t code_row_stride,
                           const float *norms_or_null, int mem_kind, void *stream, rb_qstore **out);
void rb_qstore_destroy(rb_qstore *store);
size_t rb_qstore_len(const rb_qstore *store);
int rb_qstore_has_norms(const rb_qstore *store);
/* out[i, :] = reconstruct(codes[indices[i]]) (projection applied as in pq.rs:323-326) * norms[indices[i]]:
 * bit-exact without a projection, within 1e-5 with one (as rb_pq_reconstruct_batch).  indices: [n_idx] u64 and out:
 * [n_idx, d] with element strides, both in mem_kind memory.  An index >= len answers RB_ERR_INVALID (the reference
 * panics on the out-of-bounds row).  Synchronises `stream`. */
rb_status rb_qstore_embeddings(const rb_qstore *store, const uint64_t *indices, size_t n_idx, float *out,
                               ptrdiff_t out_row_stride, ptrdiff_t out_col_stride, int mem_kind, void *stream);
/* Fused decode + dot: out[q, i] = queries[q, :] . embedding(i) for every stored row, without materialising the
 * [n, d] reconstruction: per query one table of subvector-centroid dot products (the query is rotated by the
 * projection first, q . (y R^T) = (q R) . y), then one pass over the codes per batch of up to 8 queries (as many as fit their lookup table in shared memory: 4 at M = 30, k = 256).  f32
 * accumulation, subquantizers in ascending order: within 1e-5 * sum_c |q_c| |e_c| of the exact product.
 * queries: [nq, d] with element strides; out: [nq, len] with row stride out_row_stride; both in mem_kind memory.
 * DEVICE calls are asynchronous on `stream`. */
rb_status rb_qstore_dot(const rb_qstore *store, const float *queries, size_t nq, ptrdiff_t query_row_stride,
                        ptrdiff_t query_col_stride, float *out, ptrdiff_t out_row_stride, int mem_kind, void *stream);

/* ---- A = f64 --------------------------------------------------------------------------------------- */

/* Pq<A> is generic over NdFloat (pq.rs:29-32,196-203; linalg.rs:150-156); every configuration this library was built
 * for and every known caller (finalfusion's quantized embeddings) uses A = f32, and the kernels here are f32 only.
 * A binding routes its f64 instantiation to these entry points, which all return RB_ERR_UNSUPPORTED with a message
 * (no silent down-conversion: f32 arithmetic cannot reproduce the reference's f64 codes bit for bit).  The binding
 * keeps calling the reference's own CPU code for f64. */
rb_status rb_pq_create_f64(const double *quantizers, size_t n_subquantizers, size_t n_centroids,
                           size_t subquantizer_dim, const double *projection_or_null, rb_pq **out);
rb_status rb_pq_quantize_batch_f64(const rb_pq *pq, const double *x, size_t n, ptrdiff_t x_row_stride,
                                   ptrdiff_t x_col_stride, void *codes, int code_width, ptrdiff_t code_row_stride,
                                   ptrdiff_t code_col_stride, int mem_kind, void *stream);
rb_status rb_pq_reconstruct_batch_f64(const rb_pq *pq, const void *codes, int code_width, size_t n,
                                      ptrdiff_t code_row_stride, ptrdiff_t code_col_stride, double *out,
                                      ptrdiff_t out_row_stride, ptrdiff_t out_col_stride, int mem_kind, void *stream);
rb_status rb_pq_train_f64(const double *instances, size_t n, size_t d, ptrdiff_t row_stride, ptrdiff_t col_stride,
                          size_t n_subquantizers, uint32_t n_subquantizer_bits, size_t n_iterations,
                          size_t n_attempts, const double *initial_centroids, double *loss_out, int mem_kind,
                          void *stream, rb_pq **out);

/* ---- multi-GPU training (data-parallel Pq k-means; NCCL is called inside the library) ------------- */

/* The reference trains the M subquantizers as M independent Rayon tasks (pq.rs:226-241), each a sequential k-means
 * (kmeans.rs:263-327).  Across G GPUs the two phases of an iteration are sharded on different axes:
 *   cluster_assignments (kmeans.rs:319): by ROWS    -- rank r assigns its own rows against all M codebooks;
 *   update_centroids    (kmeans.rs:320): by SUBQUANTIZERS -- rank r owns subquantizers rb_dist_subquantizer_range(r)
 *     and adds all n rows of every cluster sequentially in row order, which is exactly the reference's f32 chain
 *     (kmeans.rs:185-189): trained centroids are BIT-IDENTICAL to a one-GPU run for any G.
 * Rank r's rows follow rank r-1's in the reference's row order.  Exchanges: the column slices of the training matrix
 * once (when the state is created), then per iteration the u8 assignments (all-to-all, n*M bytes in total) and the
 * new centroids (all-gather). */
typedef struct rb_comm rb_comm;               /* one rank of an NCCL communicator, bound to the current device */
typedef struct rb_kmeans_dist rb_kmeans_dist; /* both layouts of the training rows + exchange buffers of one run */

/* Subquantizers [*m_begin, *m_end) whose centroid update rank `rank` of `world` owns (host only). */
rb_status rb_dist_subquantizer_range(size_t n_subquantizers, int rank, int world, size_t *m_begin, size_t *m_end);

/* ncclGetUniqueId into id_out (len >= 128); rank 0 creates it and hands it to the other ranks by any means. */
rb_status rb_comm_unique_id(void *id_out, size_t len);
/* ncclCommInitRank on the CURRENT device (collective over the `world` ranks). */
rb_status rb_comm_create(const void *id, int rank, int world, rb_comm **out);
void rb_comm_destroy(rb_comm *comm);
int rb_comm_rank(const rb_comm *comm);
int rb_comm_world(const rb_comm *comm);

/* x_local: DEVICE [n_local, M*dsub] with row stride x_row_stride, borrowed for the lifetime of the state (the rows
 * must not change: k-means never modifies its instances).  Collective: exchanges the row counts and the column
 * slices of the training matrix. */
rb_status rb_kmeans_dist_create(rb_comm *comm, const float *x_local, size_t n_local, ptrdiff_t x_row_stride,
                                size_t n_subquantizers, size_t n_centroids, size_t subquantizer_dim, void *stream,
                                rb_kmeans_dist **out);
/* One KMeansIteration::kmeans_iteration (kmeans.rs:308-327) for all M subquantizers over the rows of all ranks.
 * centroids: DEVICE [M,k,dsub], identical on every rank on entry and on exit (updated in place); loss_or_null:
 * DEVICE [M].  Collective, asynchronous on `stream`. */
rb_status rb_kmeans_dist_iterate(rb_kmeans_dist *state, float *centroids, float *loss_or_null, void *stream);
void rb_kmeans_dist_destroy(rb_kmeans_dist *state);
/* 1: the assignments are stored by the assignment kernels straight into their owners' memory (the [M][pitch] code
 * matrix is one peer-mapped address range: cuMemCreate / cuMemMap, descriptors passed between the ranks of the node);
 * 0: they are exchanged with ncclSend/Recv (single rank, no peer access, RB_DIST_P2P=0).  Same result either way. */
int rb_kmeans_dist_peer_window(const rb_kmeans_dist *state);

/* TrainPq::train_pq_using for Pq (pq.rs:201-249) over rows sharded across the ranks of `comm`; every rank passes
 * its own rows (instances_local: [n_local, d], mem_kind; unit column stride) and the same initial centroids (HOST
 * [n_attempts, M, k, dsub]) and receives the same quantizer.  n_total = rows of all ranks (validated). */
rb_status rb_pq_train_dist(rb_comm *comm, const float *instances_local, size_t n_local, size_t n_total, size_t d,
                           ptrdiff_t row_stride, size_t n_subquantizers, uint32_t n_subquantizer_bits,
                           size_t n_iterations, size_t n_attempts, const float *initial_centroids, float *loss_out,
                           int mem_kind, void *stream, rb_pq **out);
/* The same from ONE process: instances HOST [n, d] are split into contiguous row blocks over `devices`
 * (ncclCommInitAll, one host thread per device).  The returned quantizer lives on devices[0]. */
rb_status rb_pq_train_multi(const int *devices, int n_devices, const float *instances, size_t n, size_t d,
                            ptrdiff_t row_stride, size_t n_subquantizers, uint32_t n_subquantizer_bits,
                            size_t n_iterations, size_t n_attempts, const float *initial_centroids, float *loss_out,
                            rb_pq **out);

#ifdef __cplusplus
}
#endif
#endif /* REDUCTIVE_B200_H */
