"""Multi-GPU layer: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch) for the plumbing.

  * quantize_batch / reconstruct_batch shard by rows with the codebook replicated: NO collective
    (SURVEY §8e).  `shard_rows` is the whole of it.
  * Data-parallel Pq k-means (BASELINE config C3): rows sharded; every iteration each rank runs
    assign + segmented sum on its rows (rb_kmeans_assign_accumulate), ONE all-reduce(sum) of the packed
    accumulator  sums [M,k,dsub] | counts [M,k] | sumsq [M]  (C3: 96*256*9*4 B + 384 B = 885 KB), then the
    identical finalize on every rank (rb_kmeans_finalize), so centroids stay replicated without a broadcast.
    It restates the loop of kmeans_with_centroids (src/kmeans.rs:263-288) around kmeans_iteration
    (src/kmeans.rs:308-327) for all M subquantizers of Pq::train_pq_using (src/pq/pq.rs:201-249) at once.

    A second mode ("chained") trades the all-reduce for a rank-to-rank relay of the running sums and is
    bit-identical to a one-process run (see kmeans_data_parallel).

`local_step` / `assign` / `accumulate` / `finalize` are injectable so that the sharding + all-reduce logic is testable on CPU with the
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


def cuda_assign(x_local, centroids):
    """cluster_assignments of this rank's rows (src/kmeans.rs:319); returns the device buffer of assignments."""
    import torch

    M, k, dsub = centroids.shape
    n = x_local.shape[0]
    nbytes = M * int(lib.rb_kmeans_code_pitch(n)) * int(lib.rb_kmeans_code_width(k))
    codes = torch.empty((nbytes,), dtype=torch.uint8, device=x_local.device)
    check(lib.rb_kmeans_assign(x_local.data_ptr(), n, x_local.stride(0), centroids.data_ptr(), M, k, dsub,
                               codes.data_ptr(), torch.cuda.current_stream().cuda_stream))
    return codes


def cuda_accumulate(x_local, centroids, codes, packed_before, packed, m0: int = 0, m1: Optional[int] = None) -> None:
    """update_centroids' scatter-add of this rank's rows for the subquantizers [m0, m1), continuing the chains of the
    preceding ranks (src/kmeans.rs:181-189).  `packed_before` / `packed` hold that slice only:
    sums [(m1-m0),k,dsub] | counts [(m1-m0),k] | sum of squared norms [(m1-m0)]."""
    import torch

    M, k, dsub = centroids.shape
    m1 = M if m1 is None else m1
    n = x_local.shape[0]
    pitch, width = int(lib.rb_kmeans_code_pitch(n)), int(lib.rb_kmeans_code_width(k))
    check(lib.rb_kmeans_accumulate(x_local.data_ptr() + 4 * m0 * dsub, n, x_local.stride(0),
                                   codes.data_ptr() + m0 * pitch * width, m1 - m0, k, dsub,
                                   None if packed_before is None else packed_before.data_ptr(), packed.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream))


def cuda_finalize(packed, n_total: int, centroids, loss) -> None:
    """divide / zero empties / loss on the all-reduced sums (src/kmeans.rs:191-197, 330-360)."""
    import torch

    M, k, dsub = centroids.shape
    check(lib.rb_kmeans_finalize(packed.data_ptr(), M, k, dsub, n_total, centroids.data_ptr(),
                                 None if loss is None else loss.data_ptr(), torch.cuda.current_stream().cuda_stream))


def slice_len(m0: int, m1: int, k: int, dsub: int) -> int:
    return (m1 - m0) * (k * dsub + k + 1)


class Comm:
    """One rank of an NCCL communicator owned by the CUDA library (rb_comm).  The 128-byte NCCL id is created on
    rank 0 and handed to the other ranks through `torch.distributed` (any backend) -- plumbing only: every exchange
    of the training loop itself is issued from C++ (csrc/dist.cu)."""

    def __init__(self, rank: Optional[int] = None, world: Optional[int] = None, group=None):
        import ctypes as C

        import torch
        import torch.distributed as dist

        if rank is None:
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = torch.zeros((128,), dtype=torch.uint8)
        if rank == 0:
            buf = (C.c_ubyte * 128)()
            check(lib.rb_comm_unique_id(buf, 128))
            ident = torch.frombuffer(bytearray(buf), dtype=torch.uint8).clone()
        if world > 1:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else None
            t = ident.to(dev) if dev is not None else ident
            dist.broadcast(t, src=dist.get_process_group_ranks(group)[0] if group is not None else 0, group=group)
            ident = t.cpu()
        self.rank, self.world = rank, world
        self._h = C.c_void_p()
        raw = (C.c_ubyte * 128).from_buffer_copy(ident.numpy().tobytes())
        check(lib.rb_comm_create(raw, rank, world, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self) -> None:
        if self._h:
            lib.rb_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass


def subquantizer_range(M: int, rank: int, world: int) -> Tuple[int, int]:
    """Subquantizers whose centroid update `rank` owns in the sharded k-means (rb_dist_subquantizer_range)."""
    import ctypes as C

    a, b = C.c_size_t(), C.c_size_t()
    check(lib.rb_dist_subquantizer_range(M, rank, world, C.byref(a), C.byref(b)))
    return int(a.value), int(b.value)


class ShardedKMeans:
    """Data-parallel Pq k-means with rows sharded over the ranks of `comm` (rb_kmeans_dist): assignment is sharded by
    rows, the ordered centroid update by subquantizers, so the result is bit-identical to a one-GPU run for any
    number of GPUs (src/kmeans.rs:185-189 adds every cluster's rows sequentially in row order).  `x_local` is this
    rank's block of rows (rank r's rows follow rank r-1's) and must stay alive and unchanged."""

    def __init__(self, comm: Comm, x_local, M: int, k: int, dsub: int):
        import ctypes as C

        import torch

        assert x_local.is_cuda and x_local.dtype == torch.float32 and x_local.stride(1) == 1
        self.comm, self.x_local, self.shape = comm, x_local, (M, k, dsub)
        self._h = C.c_void_p()
        check(lib.rb_kmeans_dist_create(comm.handle, x_local.data_ptr(), x_local.shape[0], x_local.stride(0), M, k, dsub,
                                        torch.cuda.current_stream().cuda_stream, C.byref(self._h)))

    def iterate(self, centroids, loss=None) -> None:
        """One kmeans_iteration (src/kmeans.rs:308-327) in place on `centroids` [M,k,dsub] (device, identical on
        every rank); `loss`: optional device [M]."""
        import torch

        assert tuple(centroids.shape) == self.shape and centroids.is_contiguous()
        check(lib.rb_kmeans_dist_iterate(self._h, centroids.data_ptr(), None if loss is None else loss.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream))

    @property
    def peer_window(self) -> bool:
        """True: the assignment kernels store the codes straight into their owners' memory (peer-mapped code matrix);
        False: the codes travel by ncclSend/Recv (or there is one rank)."""
        return bool(lib.rb_kmeans_dist_peer_window(self._h))

    def close(self) -> None:
        if self._h:
            import torch

            torch.cuda.synchronize()
            lib.rb_kmeans_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def kmeans_data_parallel(x_local, n_total: int, centroids, n_iterations: int, group=None,
                         local_step: Optional[Callable] = None, finalize: Optional[Callable] = None,
                         on_iteration: Optional[Callable] = None, mode: str = "allreduce",
                         assign: Optional[Callable] = None, accumulate: Optional[Callable] = None,
                         n_slices: Optional[int] = None):
    """Run `n_iterations` data-parallel Lloyd iterations IN PLACE on `centroids` ([M,k,dsub], identical on every
    rank); returns the per-subquantizer loss of the last iteration ([M] tensor).

    x_local: this rank's rows [n_local, d] (d = M*dsub, unit column stride); rank r holds the rows that follow
    rank r-1's in the global row order (shard_rows).

    mode "allreduce" (BASELINE config C3): ONE all-reduce(sum) of the packed per-rank sums per iteration.  The
      per-cluster sums are then added in a different order than the reference's single sequential f32 chain
      (src/kmeans.rs:185-189); k-means amplifies such last-bit differences (a flipped near-tie moves two
      centroids), so trained centroids agree with a one-process run only to ~1e-3 relative after ~10 iterations
      at n = 1M, although the loss agrees to ~1e-6.
    mode "chained": every rank assigns its rows in parallel, then the running sums travel rank 0 -> 1 -> ... ->
      last (point-to-point) with each rank CONTINUING the chains over its rows, and the last rank broadcasts the
      totals.  The result is bit-identical to one sequential pass — the oracle's and the reference's — for any
      number of GPUs.  The subquantizers are independent, so they travel in `n_slices` slices (default:
      min(world, 4); per-slice launches have a fixed cost) as a wavefront: while rank r continues slice s, rank r-1 already works on slice s+1, and the relay costs
      about (world + n_slices - 1) / n_slices local updates instead of `world`."""
    import torch
    import torch.distributed as dist

    local_step = local_step or cuda_local_step
    finalize = finalize or cuda_finalize
    assign = assign or cuda_assign
    accumulate = accumulate or cuda_accumulate
    if mode not in ("allreduce", "chained"):
        raise ValueError(f"unknown mode {mode!r}")
    M, k, dsub = centroids.shape
    plen = M * k * dsub + M * k + M
    loss = torch.zeros((M,), dtype=torch.float32, device=centroids.device)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if distributed and mode == "chained":
        # the relay continues chains with rb_kmeans_accumulate(packed_before): every rank checks its preconditions
        # BEFORE the first collective, so that an unsupported shape raises everywhere instead of hanging the ranks that
        # already wait in recv / broadcast (prefer ShardedKMeans: bit-identical as well, and it scales)
        if k > 256 or dsub not in (1, 2, 3, 4, 5, 6, 8, 10, 12, 15, 16, 20, 30, 32) or x_local.stride(0) >= 1 << 29:
            raise NotImplementedError(f"chained mode needs k <= 256 and an instantiated subvector width (k={k}, dsub={dsub})")
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
        ns = max(1, min(M, n_slices if n_slices is not None else min(world, 4)))
        bounds = [(M * i) // ns for i in range(ns + 1)]
        slices = [(bounds[i], bounds[i + 1]) for i in range(ns) if bounds[i + 1] > bounds[i]]
        new = lambda a, b: torch.empty((slice_len(a, b, k, dsub),), dtype=torch.float32, device=centroids.device)  # noqa: E731
        packed_s = [new(a, b) for a, b in slices]
        before_s = [new(a, b) for a, b in slices] if rank > 0 else [None] * len(slices)
    else:
        packed = torch.empty((plen,), dtype=torch.float32, device=centroids.device)
    for it in range(n_iterations):
        if distributed and mode == "chained":
            codes = assign(x_local, centroids)                       # all ranks at once
            for (a, b), before, mine in zip(slices, before_s, packed_s):
                if rank > 0:
                    dist.recv(before, src=ranks[rank - 1], group=group)  # sums over the rows of ranks < rank
                accumulate(x_local, centroids, codes, before, mine, a, b)
                if rank + 1 < world:
                    dist.send(mine, dst=ranks[rank + 1], group=group)
            for (a, b), mine in zip(slices, packed_s):
                dist.broadcast(mine, src=ranks[world - 1], group=group)
                finalize(mine, n_total, centroids[a:b], loss[a:b])
        else:
            local_step(x_local, centroids, packed)
            if distributed:
                dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)  # the one exchange per iteration
            finalize(packed, n_total, centroids, loss)
        if on_iteration is not None:
            on_iteration(it)
    return loss
