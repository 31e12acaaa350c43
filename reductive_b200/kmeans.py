"""Host-side mirror of reductive's k-means API (src/kmeans.rs) over the C ABI.

The arithmetic runs in the CUDA kernels (assignment = the encode kernel, update = segmented sum +
finalize); this module only holds device buffers (torch) and the loop structure of
  KMeansIteration::kmeans_iteration        src/kmeans.rs:308-327
  KMeansWithCentroids::kmeans_with_centroids  src/kmeans.rs:263-288
  KMeans::k_means                           src/kmeans.rs:224-239
for instances along Axis(0) (the only axis the PQ path uses, pq.rs:177, opq.rs:207; an Axis(1) caller passes
the transposed view, which is a stride swap at this boundary).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from ._cabi import ReductivePanic, check, lib


class NIterationsCondition:
    """Condition that stops clustering after N iterations (src/kmeans.rs:97-104)."""

    def __init__(self, n: int):
        self.n = n

    def should_stop(self, iteration: int, loss: float) -> bool:
        return iteration >= self.n


class RandomInstanceCentroids:
    """Pick random data set instances as centroids (src/kmeans.rs:35-88)."""

    def __init__(self, rng: np.random.Generator):
        self.rng = rng

    def initial_centroids(self, data, k: int):
        n = data.shape[0]
        if k <= 0:
            raise ReductivePanic("Cannot pick 0 random centroids")
        if k >= n:
            raise ReductivePanic(f"Cannot pick more centroids than instances: {n} instances, {k} centroids")
        if data.shape[1] == 0:
            raise ReductivePanic("Cannot pick centroids from zero-length instances")
        idx = self.rng.choice(n, size=k, replace=False)
        return _to_device(data)[_index(idx, data)].clone()


def _torch():
    import torch

    if not torch.cuda.is_available():
        from ._cabi import NoDeviceError

        raise NoDeviceError("no CUDA device visible; reductive_b200 has no CPU fallback")
    return torch


def _to_device(a):
    torch = _torch()
    if isinstance(a, torch.Tensor):
        return a if a.is_cuda else a.cuda()
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


def _index(idx, like):
    torch = _torch()
    return torch.as_tensor(np.asarray(idx), device="cuda")


def kmeans_iteration(instances, centroids) -> float:
    """One Lloyd iteration IN PLACE on `centroids` (a CUDA tensor [k, dsub]); returns the mean squared error
    (KMeansIteration::kmeans_iteration, src/kmeans.rs:308-327).  `instances`: CUDA tensor [n, dsub], unit
    column stride."""
    torch = _torch()
    x = instances
    if centroids.shape[0] == 0:
        raise ReductivePanic("Cannot cluster instances with zero centroids.")  # kmeans.rs:309-312
    if centroids.shape[1] != x.shape[1]:
        raise ReductivePanic("Centroid and instance lengths differ.")  # kmeans.rs:313-317
    if x.stride(1) != 1:
        x = x.contiguous()
    k, dsub = centroids.shape
    assert centroids.is_contiguous() and centroids.dtype == torch.float32 and x.dtype == torch.float32
    stream = torch.cuda.current_stream().cuda_stream
    packed = torch.empty((lib.rb_kmeans_packed_len(1, k, dsub),), dtype=torch.float32, device=x.device)
    loss = torch.empty((1,), dtype=torch.float32, device=x.device)
    check(lib.rb_kmeans_assign_accumulate(x.data_ptr(), x.shape[0], x.stride(0), centroids.data_ptr(), 1, k, dsub,
                                          packed.data_ptr(), stream))
    check(lib.rb_kmeans_finalize(packed.data_ptr(), 1, k, dsub, x.shape[0], centroids.data_ptr(), loss.data_ptr(),
                                 stream))
    return float(loss.item())


def kmeans_with_centroids(instances, centroids, stop_condition) -> float:
    """KMeansWithCentroids::kmeans_with_centroids (src/kmeans.rs:263-288): iterate until the stop condition."""
    it = 0
    while True:
        loss = kmeans_iteration(instances, centroids)
        it += 1
        if stop_condition.should_stop(it, loss):
            return loss


class KMeans:
    """KMeans::k_means (src/kmeans.rs:224-239)."""

    @staticmethod
    def k_means(instances, k: int, initial_centroids, stop_condition) -> Tuple[np.ndarray, float]:
        x = _to_device(instances)
        if k == 0 or k > x.shape[0]:
            raise ReductivePanic("k cannot be larger than the number of data points or zero")
        cen = initial_centroids.initial_centroids(x, k).contiguous()
        loss = kmeans_with_centroids(x, cen, stop_condition)
        return cen.cpu().numpy(), loss
