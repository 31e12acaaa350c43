"""Opq / GaussianOpq training orchestration (SURVEY §8f "next" row) over the same CUDA kernels.

Mirror of  Opq          src/pq/opq.rs:38-209  (Ge et al., 2013; alternating k-means step / Procrustes rotation)
           GaussianOpq  src/pq/gaussian_opq.rs:25-68  (PCA + eigenvalue bucketing once, then plain Pq training)
           bucket_eigenvalues  src/pq/opq.rs:212-273
Both RETURN A `Pq` carrying the projection, so encode/decode always run through Pq (x.R before the argmin,
pq.rs:276; R^T after the gather, pq.rs:323-326).

What runs where: everything over the n rows runs in this repo's kernels through the C ABI -- the covariance
(rb_covariance), and per training iteration X.R, the k-means step, the quantize -> reconstruct round trip and the
Gram matrix X^T.Y^ (rb_opq_train_iteration).  Only the d x d eigendecomposition and SVD stay on host LAPACK (numpy),
exactly as the reference does (opq.rs:123,187) and north_star scopes it.  torch is used for device memory only.
"""
from __future__ import annotations

from typing import List

import numpy as np

from ._cabi import ReductivePanic, check, lib
from .pq import Pq, TrainPq, check_quantizer_invariants


def bucket_eigenvalues(eigenvalues, n_buckets: int) -> List[List[int]]:
    """Distribute principal directions over buckets so that the products of eigenvalues are balanced
    (src/pq/opq.rs:212-273)."""
    ev = np.asarray(eigenvalues)
    if ev.dtype not in (np.float32, np.float64):
        ev = ev.astype(np.float64)
    if n_buckets <= 0:
        raise ReductivePanic("Cannot distribute eigenvalues over zero buckets.")
    if len(ev) < n_buckets:
        raise ReductivePanic("At least one eigenvalue is required per bucket")
    if len(ev) % n_buckets != 0:
        raise ReductivePanic("The number of eigenvalues should be a multiple of the number of buckets.")
    eps = np.finfo(ev.dtype).eps
    order = sorted(range(len(ev)), key=lambda i: (np.isnan(ev[i]), ev[i]))  # OrderedFloat ascending
    if not ev[order[0]] >= -eps:
        raise ReductivePanic("Bucketing is only supported for positive eigenvalues.")
    logs = np.log(ev + eps).astype(ev.dtype)  # log-space products, opq.rs:244
    logs = logs - logs.min()                   # opq.rs:248-253
    assignments: List[List[int]] = [[] for _ in range(n_buckets)]
    products = [ev.dtype.type(0)] * n_buckets
    max_assignments = len(ev) // n_buckets
    while order:
        idx_ev = order.pop()  # largest remaining eigenvalue
        best = None
        for b in range(n_buckets):  # non-full bucket with the smallest product, first on ties (opq.rs:261-266)
            if len(assignments[b]) < max_assignments and (best is None or products[b] < products[best]):
                best = b
        assignments[best].append(idx_ev)
        products[best] = ev.dtype.type(products[best] + logs[idx_ev])
    return assignments


def _torch():
    import torch

    if not torch.cuda.is_available():
        from ._cabi import NoDeviceError

        raise NoDeviceError("no CUDA device visible; reductive_b200 has no CPU fallback")
    return torch


def _device_rows(instances):
    torch = _torch()
    if isinstance(instances, torch.Tensor):
        x = instances if instances.is_cuda else instances.cuda()
    else:
        x = torch.from_numpy(np.ascontiguousarray(instances, np.float32)).cuda()
    return x.float()


def create_projection_matrix(x, n_subquantizers: int) -> np.ndarray:
    """Opq::create_projection_matrix (src/pq/opq.rs:103-136): covariance (src/linalg.rs:23-44) -> eigh ->
    bucket the eigenvalues -> eigenvectors permuted bucket by bucket as columns."""
    torch = _torch()
    n = x.shape[0]
    if n == 0:
        raise ReductivePanic("Cannot compute a covariance from zero observations")
    cov_dev = torch.empty((x.shape[1], x.shape[1]), dtype=torch.float32, device=x.device)
    check(lib.rb_covariance(x.data_ptr(), n, x.shape[1], x.stride(0), cov_dev.data_ptr(),
                            torch.cuda.current_stream().cuda_stream))
    cov = cov_dev.cpu().numpy()
    eigen_values, eigen_vectors = np.linalg.eigh(cov, UPLO="U")  # host LAPACK, as in the reference
    buckets = bucket_eigenvalues(eigen_values, n_subquantizers)
    order = [i for b in buckets for i in b]
    return np.ascontiguousarray(eigen_vectors[:, order], np.float32)


def _project(x, r_dev, transpose: bool = False):
    torch = _torch()
    out = torch.empty((x.shape[0], x.shape[1]), dtype=torch.float32, device=x.device)
    check(lib.rb_project_rows(x.data_ptr(), x.shape[0], x.shape[1], x.stride(0), x.stride(1), r_dev.data_ptr(),
                              1 if transpose else 0, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return out


class Opq(TrainPq):
    """Optimized product quantizer (Ge et al., 2013).  `n_attempts` has no effect (opq.rs:35-37)."""

    @classmethod
    def train_pq_using(cls, n_subquantizers, n_subquantizer_bits, n_iterations, _n_attempts, instances, rng):
        torch = _torch()
        x = _device_rows(instances)
        n, d = x.shape
        check_quantizer_invariants(n_subquantizers, n_subquantizer_bits, n_iterations, 1, n, d)  # opq.rs:58-64
        M, k = n_subquantizers, 1 << n_subquantizer_bits
        dsub = d // M
        projection = torch.from_numpy(create_projection_matrix(x, M)).cuda()  # opq.rs:67
        rx = _project(x, projection)  # opq.rs:68
        if k >= n:
            raise ReductivePanic(f"Cannot pick more centroids than instances: {n} instances, {k} centroids")
        # opq.rs:71-76,138-159: one rng, sequentially per subquantizer
        cen = torch.empty((M, k, dsub), dtype=torch.float32, device=x.device)
        for m in range(M):
            idx = torch.as_tensor(rng.choice(n, size=k, replace=False), device=x.device)
            cen[m] = rx[idx, m * dsub:(m + 1) * dsub]
        stream = torch.cuda.current_stream().cuda_stream
        xty = torch.empty((d, d), dtype=torch.float32, device=x.device)
        for _ in range(n_iterations):  # opq.rs:86-93 -> train_iteration opq.rs:161-189
            # device part: rx = X.R, one k-means step per subquantizer, quantize -> reconstruct, X^T.Y^
            check(lib.rb_opq_train_iteration(x.data_ptr(), n, d, x.stride(0), projection.data_ptr(), cen.data_ptr(), M, k,
                                             xty.data_ptr(), stream))
            # Procrustes: R = U V^T of X^T Y^ (opq.rs:187-188); d x d SVD on host LAPACK
            u, _, vt = np.linalg.svd(xty.cpu().numpy(), full_matrices=True)
            projection = torch.from_numpy(np.ascontiguousarray(u @ vt, np.float32)).cuda()
        return Pq(projection.cpu().numpy(), cen.cpu().numpy())


class GaussianOpq(TrainPq):
    """Optimized product quantizer for Gaussian variables (src/pq/gaussian_opq.rs:25-68)."""

    @classmethod
    def train_pq_using(cls, n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, instances, rng):
        torch = _torch()
        x = _device_rows(instances)
        n, d = x.shape
        check_quantizer_invariants(n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, n, d)
        projection = create_projection_matrix(x, n_subquantizers)  # gaussian_opq.rs:53
        rx = _project(x, torch.from_numpy(projection).cuda())   # gaussian_opq.rs:54
        pq = Pq.train_pq_using(n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, rx, rng)
        return Pq(projection, pq.subquantizers())  # gaussian_opq.rs:64-67
