"""ctypes loader for the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing in reductive_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

ERR_NAMES = {
    0: "Ok",
    1: "IncorrectNAttempts",
    2: "IncorrectNIterations",
    3: "IncorrectNSubquantizerBits",
    4: "IncorrectNumberSubquantizers",
    5: "NSubquantizersOutsideRange",
}


def build(force: bool = False) -> None:
    """Compile liboracle.so / liboracle_scalar.so with the committed Makefile."""
    targets = ["liboracle.so", "liboracle_scalar.so"]
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("oracle.c", "oracle.h", "Makefile"))
    need = force or any(
        not os.path.exists(os.path.join(_HERE, t)) or os.path.getmtime(os.path.join(_HERE, t)) < src_m
        for t in targets
    )
    if need:
        subprocess.run(["make", "-C", _HERE] + targets, check=True, capture_output=True)


_fp = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)


def _f(a):
    return a.ctypes.data_as(_fp)


class Oracle:
    """Thin typed wrapper over one built oracle library."""

    def __init__(self, scalar: bool = False):
        build()
        name = "liboracle_scalar.so" if scalar else "liboracle.so"
        self.lib = lib = C.CDLL(os.path.join(_HERE, name))
        sz, pd = C.c_size_t, C.c_ssize_t
        lib.orc_unrolled_dot.restype = C.c_float
        lib.orc_unrolled_dot.argtypes = [_fp, _fp, sz]
        lib.orc_strided_dot.restype = C.c_float
        lib.orc_strided_dot.argtypes = [_fp, pd, _fp, pd, sz]
        lib.orc_sgemm.restype = None
        lib.orc_sgemm.argtypes = [sz, sz, sz, _fp, pd, pd, _fp, pd, pd, _fp, pd, pd]
        lib.orc_sqdist_batch.restype = None
        lib.orc_sqdist_batch.argtypes = [_fp, sz, sz, _fp, sz, sz, _fp]
        lib.orc_sqdist_vec.restype = None
        lib.orc_sqdist_vec.argtypes = [_fp, _fp, sz, sz, _fp]
        lib.orc_cluster_assignments.restype = None
        lib.orc_cluster_assignments.argtypes = [_fp, sz, sz, _fp, sz, sz, _u64p]
        lib.orc_cluster_assignment.restype = C.c_uint64
        lib.orc_cluster_assignment.argtypes = [_fp, sz, sz, _fp]
        lib.orc_update_centroids.restype = None
        lib.orc_update_centroids.argtypes = [_fp, sz, sz, _fp, sz, sz, _u64p]
        lib.orc_mean_squared_error.restype = C.c_float
        lib.orc_mean_squared_error.argtypes = [_fp, sz, sz, _fp, sz, sz, _u64p]
        lib.orc_kmeans_iteration.restype = C.c_float
        lib.orc_kmeans_iteration.argtypes = [_fp, sz, sz, _fp, sz, sz]
        lib.orc_kmeans_with_centroids.restype = C.c_float
        lib.orc_kmeans_with_centroids.argtypes = [_fp, sz, sz, _fp, sz, sz, sz]
        lib.orc_check_quantizer_invariants.restype = C.c_int
        lib.orc_check_quantizer_invariants.argtypes = [sz, C.c_uint32, sz, sz, sz, sz, _u64p]
        lib.orc_quantize_batch.restype = None
        lib.orc_quantize_batch.argtypes = [_fp, sz, sz, sz, _fp, _fp, sz, pd, pd, C.c_void_p, C.c_int, pd, pd, C.c_int]
        lib.orc_quantize_vector.restype = C.c_int
        lib.orc_quantize_vector.argtypes = [_fp, sz, sz, sz, _fp, _fp, pd, C.c_void_p, C.c_int]
        lib.orc_reconstruct_batch.restype = C.c_int
        lib.orc_reconstruct_batch.argtypes = [_fp, sz, sz, sz, _fp, C.c_void_p, C.c_int, sz, pd, pd, _fp, pd, pd, C.c_int]
        lib.orc_reconstruct.restype = C.c_int
        lib.orc_reconstruct.argtypes = [_fp, sz, sz, sz, _fp, C.c_void_p, C.c_int, _fp]
        lib.orc_train_pq.restype = C.c_int
        lib.orc_train_pq.argtypes = [_fp, sz, sz, sz, C.c_uint32, sz, sz, _fp, _fp, _fp, C.c_int]
        lib.orc_has_avx2_kernel.restype = C.c_int

    # ---- linalg -------------------------------------------------------------------------
    def unrolled_dot(self, x, y):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(y, np.float32)
        return np.float32(self.lib.orc_unrolled_dot(_f(x), _f(y), x.size))

    def sgemm(self, a, b):
        """a.dot(b) with the strides of the numpy views honoured (matrixmultiply model)."""
        assert a.dtype == np.float32 and b.dtype == np.float32
        m, k = a.shape
        k2, n = b.shape
        assert k == k2
        c = np.zeros((m, n), np.float32)
        es = 4
        self.lib.orc_sgemm(m, k, n, _f(a), a.strides[0] // es, a.strides[1] // es,
                           _f(b), b.strides[0] // es, b.strides[1] // es, _f(c), n, 1)
        return c

    def sqdist_batch(self, x, c):
        x = np.ascontiguousarray(x, np.float32)
        c = np.ascontiguousarray(c, np.float32)
        out = np.zeros((x.shape[0], c.shape[0]), np.float32)
        self.lib.orc_sqdist_batch(_f(x), x.shape[0], x.shape[1], _f(c), c.shape[0], c.shape[1], _f(out))
        return out

    def sqdist_vec(self, x, c):
        x = np.ascontiguousarray(x, np.float32)
        c = np.ascontiguousarray(c, np.float32)
        out = np.zeros((c.shape[0],), np.float32)
        self.lib.orc_sqdist_vec(_f(x), _f(c), c.shape[0], c.shape[1], _f(out))
        return out

    # ---- kmeans -------------------------------------------------------------------------
    def cluster_assignments(self, centroids, instances):
        c = np.ascontiguousarray(centroids, np.float32)
        x = np.ascontiguousarray(instances, np.float32)
        out = np.zeros((x.shape[0],), np.uint64)
        self.lib.orc_cluster_assignments(_f(x), x.shape[0], x.shape[1], _f(c), c.shape[0], c.shape[1],
                                         out.ctypes.data_as(_u64p))
        return out

    def cluster_assignment(self, centroids, instance):
        c = np.ascontiguousarray(centroids, np.float32)
        x = np.ascontiguousarray(instance, np.float32)
        return int(self.lib.orc_cluster_assignment(_f(c), c.shape[0], c.shape[1], _f(x)))

    def update_centroids(self, centroids, instances, assignments):
        c = np.ascontiguousarray(centroids, np.float32).copy()
        x = np.ascontiguousarray(instances, np.float32)
        a = np.ascontiguousarray(assignments, np.uint64)
        self.lib.orc_update_centroids(_f(c), c.shape[0], c.shape[1], _f(x), x.shape[0], x.shape[1],
                                      a.ctypes.data_as(_u64p))
        return c

    def mean_squared_error(self, centroids, instances, assignments):
        c = np.ascontiguousarray(centroids, np.float32)
        x = np.ascontiguousarray(instances, np.float32)
        a = np.ascontiguousarray(assignments, np.uint64)
        return np.float32(self.lib.orc_mean_squared_error(_f(c), c.shape[0], c.shape[1], _f(x), x.shape[0],
                                                          x.shape[1], a.ctypes.data_as(_u64p)))

    def kmeans_with_centroids(self, instances, centroids, n_iterations):
        """Returns (centroids, loss).  instances may be a column slice (row stride honoured)."""
        x = instances
        assert x.dtype == np.float32 and x.strides[1] == 4
        c = np.ascontiguousarray(centroids, np.float32).copy()
        loss = self.lib.orc_kmeans_with_centroids(_f(c), c.shape[0], c.shape[1], _f(x), x.shape[0],
                                                  x.strides[0] // 4, n_iterations)
        return c, np.float32(loss)

    def kmeans_iteration(self, instances, centroids):
        return self.kmeans_with_centroids(instances, centroids, 1)

    # ---- pq -----------------------------------------------------------------------------
    def check_quantizer_invariants(self, n_subquantizers, n_bits, n_iterations, n_attempts, n_rows, n_cols):
        detail = C.c_uint64(0)
        rc = self.lib.orc_check_quantizer_invariants(n_subquantizers, n_bits, n_iterations, n_attempts,
                                                     n_rows, n_cols, C.byref(detail))
        return rc, int(detail.value)

    @staticmethod
    def _qp(quantizers, projection):
        q = np.ascontiguousarray(quantizers, np.float32)
        assert q.ndim == 3
        p = None if projection is None else np.ascontiguousarray(projection, np.float32)
        return q, p

    def quantize_batch(self, quantizers, projection, x, code_dtype=np.uint8, n_threads=1):
        q, p = self._qp(quantizers, projection)
        M, k, dsub = q.shape
        assert x.dtype == np.float32 and x.ndim == 2 and x.shape[1] == M * dsub
        codes = np.zeros((x.shape[0], M), code_dtype)
        self.lib.orc_quantize_batch(_f(q), M, k, dsub, None if p is None else _f(p),
                                    _f(x), x.shape[0], x.strides[0] // 4, x.strides[1] // 4,
                                    codes.ctypes.data_as(C.c_void_p), codes.itemsize, M, 1, n_threads)
        return codes

    def quantize_vector(self, quantizers, projection, x, code_dtype=np.uint8):
        q, p = self._qp(quantizers, projection)
        M, k, dsub = q.shape
        assert x.dtype == np.float32 and x.ndim == 1 and x.shape[0] == M * dsub
        codes = np.zeros((M,), code_dtype)
        rc = self.lib.orc_quantize_vector(_f(q), M, k, dsub, None if p is None else _f(p), _f(x),
                                          x.strides[0] // 4, codes.ctypes.data_as(C.c_void_p), codes.itemsize)
        if rc != 0:
            raise OverflowError("Cannot store centroids in quantizer index type")
        return codes

    def reconstruct_batch(self, quantizers, projection, codes, n_threads=1):
        q, p = self._qp(quantizers, projection)
        M, k, dsub = q.shape
        codes = np.ascontiguousarray(codes)
        assert codes.ndim == 2 and codes.shape[1] == M and codes.itemsize in (1, 2, 4, 8)
        out = np.zeros((codes.shape[0], M * dsub), np.float32)
        rc = self.lib.orc_reconstruct_batch(_f(q), M, k, dsub, None if p is None else _f(p),
                                            codes.ctypes.data_as(C.c_void_p), codes.itemsize, codes.shape[0],
                                            M, 1, _f(out), M * dsub, 1, n_threads)
        if rc != 0:
            raise IndexError("code out of range")
        return out

    def reconstruct(self, quantizers, projection, codes):
        q, p = self._qp(quantizers, projection)
        M, k, dsub = q.shape
        codes = np.ascontiguousarray(codes)
        out = np.zeros((M * dsub,), np.float32)
        rc = self.lib.orc_reconstruct(_f(q), M, k, dsub, None if p is None else _f(p),
                                      codes.ctypes.data_as(C.c_void_p), codes.itemsize, _f(out))
        if rc != 0:
            raise IndexError("code out of range")
        return out

    def train_pq(self, x, n_subquantizers, n_bits, n_iterations, n_attempts, initial, n_threads=1):
        """initial: [n_attempts, M, k, dsub].  Returns (quantizers [M,k,dsub], loss [M])."""
        x = np.ascontiguousarray(x, np.float32)
        n, d = x.shape
        rc, _ = self.check_quantizer_invariants(n_subquantizers, n_bits, n_iterations, n_attempts, n, d)
        if rc != 0:
            raise ValueError(ERR_NAMES[rc])
        M, k = n_subquantizers, 1 << n_bits
        init = np.ascontiguousarray(initial, np.float32).reshape(n_attempts, M, k, d // M)
        out = np.zeros((M, k, d // M), np.float32)
        loss = np.zeros((M,), np.float32)
        rc = self.lib.orc_train_pq(_f(x), n, d, M, n_bits, n_iterations, n_attempts, _f(init), _f(out),
                                   _f(loss), n_threads)
        assert rc == 0
        return out, loss


    # ---- OPQ training (test infrastructure for the "next" rows f1 / f3) ----------------------------------
    # Composed from the C primitives above (sgemm, kmeans_iteration, quantize / reconstruct) in the reference's order;
    # the d x d eigendecomposition / SVD go to host LAPACK through numpy as the reference's go through ndarray-linalg
    # (opq.rs:123, opq.rs:187), so the factors agree with the reference's only up to LAPACK driver differences.
    def covariance(self, x):
        """Covariance::covariance, observation axis 0 (linalg.rs:23-44): mean_axis (rows added in row order, then
        divided by n), centred copy, centred^T . (centred / (n - 1)) as one sgemm."""
        x = np.asarray(x, np.float32)
        n = x.shape[0]
        assert n != 0, "Cannot compute a covariance from zero observations"
        total = np.zeros((x.shape[1],), np.float32)
        for row in x:  # ndarray sum_axis(Axis(0)): lane-wise running sum over the rows
            total = total + row
        means = total / np.float32(n)
        centered = (x - means).astype(np.float32)
        scaled = (centered / (np.float32(n) - np.float32(1))).astype(np.float32)
        return self.sgemm(centered.T, scaled)

    @staticmethod
    def bucket_eigenvalues(eigenvalues, n_buckets):
        """bucket_eigenvalues (opq.rs:212-273), written independently of the product's copy: largest eigenvalue first,
        each into the non-full bucket with the smallest running sum of shifted logs (first bucket on ties)."""
        ev = np.asarray(eigenvalues)
        if ev.dtype not in (np.float32, np.float64):
            ev = ev.astype(np.float64)
        assert n_buckets > 0 and len(ev) >= n_buckets and len(ev) % n_buckets == 0
        eps = np.finfo(ev.dtype).eps
        idx = np.argsort(ev, kind="stable")  # ascending; ties keep index order like sort_unstable_by on distinct values
        assert ev[idx[0]] >= -eps, "Bucketing is only supported for positive eigenvalues."
        logs = np.log(ev + eps).astype(ev.dtype)
        logs = (logs - logs.min()).astype(ev.dtype)
        cap = len(ev) // n_buckets
        buckets = [[] for _ in range(n_buckets)]
        sums = np.zeros((n_buckets,), ev.dtype)
        for i in idx[::-1]:
            open_ = [b for b in range(n_buckets) if len(buckets[b]) < cap]
            b = min(open_, key=lambda j: (sums[j], j))
            buckets[b].append(int(i))
            sums[b] = sums[b] + logs[i]
        return buckets

    def create_projection_matrix(self, x, n_subquantizers):
        """Opq::create_projection_matrix (opq.rs:103-136)."""
        cov = self.covariance(x)
        values, vectors = np.linalg.eigh(cov, UPLO="U")
        order = [i for b in self.bucket_eigenvalues(values, n_subquantizers) for i in b]
        return np.ascontiguousarray(vectors[:, order], np.float32)

    def opq_train_iteration(self, projection, centroids, x):
        """Opq::train_iteration (opq.rs:161-189).  Returns (new projection, new centroids, X^T . Y^)."""
        x = np.ascontiguousarray(x, np.float32)
        r = np.ascontiguousarray(projection, np.float32)
        c = np.array(centroids, np.float32, copy=True)
        M, k, dsub = c.shape
        rx = self.sgemm(x, r)                                               # opq.rs:173
        for m in range(M):                                                  # opq.rs:174,191-209
            c[m], _ = self.kmeans_iteration(np.ascontiguousarray(rx[:, m * dsub:(m + 1) * dsub]), c[m])
        codes = self.quantize_batch(c, None, rx, np.uint32)                 # opq.rs:180
        reconstructed = self.reconstruct_batch(c, None, codes)              # opq.rs:181-182
        xty = self.sgemm(x.T, reconstructed)                                # opq.rs:187
        u, _, vt = np.linalg.svd(xty, full_matrices=True)
        return self.sgemm(np.ascontiguousarray(u, np.float32), np.ascontiguousarray(vt, np.float32)), c, xty

    # ---- caller-side quantized storage.  PARITY UNPINNED: the storage type lives in the finalfusion
    # crate, outside /root/reference, so no reference-held vector exists for it; these restate its documented behaviour
    # (embedding = Reconstruct::reconstruct (traits.rs:102-156, pq.rs:303-347) times the stored norm) on top of the
    # pinned reconstruct_batch above.
    def qstore_embeddings(self, quantizers, projection, codes, norms, indices):
        idx = np.asarray(indices, np.int64)
        e = self.reconstruct_batch(quantizers, projection, np.ascontiguousarray(np.asarray(codes)[idx]))
        if norms is not None:
            e = e * np.asarray(norms, np.float32)[idx][:, None]  # f32 `*=`, one rounded multiply per element
        return e.astype(np.float32)

    def qstore_dot(self, quantizers, projection, codes, norms, queries):
        """Exact (float64) scores of every row against every query and the magnitude sum_c |q_c| |e_c| the f32
        tolerance of the fused kernel is stated against."""
        e = self.qstore_embeddings(quantizers, projection, codes, norms, np.arange(len(codes))).astype(np.float64)
        q = np.asarray(queries, np.float64)
        return q @ e.T, np.abs(q) @ np.abs(e).T


_default = None


def get(scalar: bool = False) -> Oracle:
    global _default
    if scalar:
        return Oracle(scalar=True)
    if _default is None:
        _default = Oracle()
    return _default
