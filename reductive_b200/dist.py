"""Multi-GPU layer: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the plumbing.

  * quantize_batch / reconstruct_batch shard by rows with the codebook replicated: NO collective
    (SURVEY §8e).  `shard_rows` is the whole of it.
  * Data-parallel Pq k-means (BASELINE config C3): rows sharded; every iteration each rank runs
    assign + segmented sum on its rows (rb_kmeans_assign_accumulate), ONE all-reduce(sum) of the packed
    accumulator  sums [M,k,dsub] | counts [M,k] | sumsq [M]  (C3: 96*256*9*4 B + 384 B = 885 KB), then the
    identical finalize on every rank (rb_kmeans_finalize), so centroids stay replicated without a broadcast.
    It restates the loop of kmeans_with_centroids (src/kmeans.rs:263-288) around kmeans_iteration
    (src/kmeans.rs:308-327) for all M subquantizers of Pq::train_pq_using (src/pq/pq.rs:201-249) at once.

`local_step` / `finalize` are injectable so that the sharding + all-reduce logic is testable on CPU with the
gloo backend (tests/test_dist_gloo.py); the defaults call the CUDA library and fail without a GPU.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

from ._cabi import check, lib


def shard_rows(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous row block [start, stop) of `rank`: ceil(n / world_size) rows per rank (last may be short)."""
    per = -(-n // world_size)
    start = min(n, rank * per)
    return start, min(n, start + per)


def packed_len(M: int, k: int, dsub: int) -> int:
    return int(lib.rb_kmeans_packed_len(M, k, dsub))


def cuda_local_step(x_local, centroids, packed) -> None:
    """assign + accumulate on this rank's rows (src/kmeans.rs:319-325, first half)."""
    import torch

    M, k, dsub = centroids.shape
    check(lib.rb_kmeans_assign_accumulate(x_local.data_ptr(), x_local.shape[0], x_local.stride(0),
                                          centroids.data_ptr(), M, k, dsub, packed.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))


def cuda_finalize(packed, n_total: int, centroids, loss) -> None:
    """divide / zero empties / loss on the all-reduced sums (src/kmeans.rs:191-197, 330-360)."""
    import torch

    M, k, dsub = centroids.shape
    check(lib.rb_kmeans_finalize(packed.data_ptr(), M, k, dsub, n_total, centroids.data_ptr(),
                                 None if loss is None else loss.data_ptr(), torch.cuda.current_stream().cuda_stream))


def kmeans_data_parallel(x_local, n_total: int, centroids, n_iterations: int, group=None,
                         local_step: Optional[Callable] = None, finalize: Optional[Callable] = None,
                         on_iteration: Optional[Callable] = None):
    """Run `n_iterations` data-parallel Lloyd iterations IN PLACE on `centroids` ([M,k,dsub], identical on every
    rank); returns the per-subquantizer loss of the last iteration ([M] tensor).

    x_local: this rank's rows [n_local, d] (d = M*dsub, unit column stride)."""
    import torch
    import torch.distributed as dist

    local_step = local_step or cuda_local_step
    finalize = finalize or cuda_finalize
    M, k, dsub = centroids.shape
    packed = torch.empty((M * k * dsub + M * k + M,), dtype=torch.float32, device=centroids.device)
    loss = torch.zeros((M,), dtype=torch.float32, device=centroids.device)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    for it in range(n_iterations):
        local_step(x_local, centroids, packed)
        if distributed:
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)  # the one exchange per iteration
        finalize(packed, n_total, centroids, loss)
        if on_iteration is not None:
            on_iteration(it)
    return loss
