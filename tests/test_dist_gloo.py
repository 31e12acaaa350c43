"""world_size-2 gloo test of the data-parallel k-means host logic (sharding + one all-reduce per iteration +
replicated finalize).  The per-rank compute is injected as a numpy restatement of the accumulate/finalize
kernels so the test runs without a GPU; the CUDA versions are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import normal, rows_as_initial_centroids


def _np_local_step(o):
    def step(x_local, centroids, packed):
        M, k, dsub = centroids.shape
        x = x_local.numpy()
        c = centroids.numpy()
        sums = np.zeros((M, k, dsub), np.float64)
        counts = np.zeros((M, k), np.float64)
        sq = np.zeros((M,), np.float64)
        for m in range(M):
            sub = np.ascontiguousarray(x[:, m * dsub:(m + 1) * dsub])
            a = o.cluster_assignments(c[m], sub).astype(np.int64)
            np.add.at(sums[m], a, sub.astype(np.float64))
            np.add.at(counts[m], a, 1.0)
            sq[m] = (sub.astype(np.float64) ** 2).sum()
        flat = np.concatenate([sums.ravel(), counts.ravel(), sq]).astype(np.float32)
        packed.copy_(torch.from_numpy(flat))
    return step


def _np_finalize(packed, n_total, centroids, loss):
    M, k, dsub = centroids.shape
    p = packed.numpy().astype(np.float64)
    sums = p[:M * k * dsub].reshape(M, k, dsub)
    counts = p[M * k * dsub:M * k * dsub + M * k].reshape(M, k)
    sq = p[M * k * dsub + M * k:]
    c = np.where(counts[..., None] > 0, sums / np.maximum(counts[..., None], 1), 0.0)
    centroids.copy_(torch.from_numpy(c.astype(np.float32)))
    if loss is not None:
        sse = sq + (counts[..., None] * c * c - 2 * c * sums).sum(axis=(1, 2))
        loss.copy_(torch.from_numpy((sse / (n_total * dsub)).astype(np.float32)))


def _worker(rank, world, port, n, M, k, dsub, iters, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from reductive_b200.dist import kmeans_data_parallel, shard_rows

    o = orc.get()
    x = normal((n, M * dsub), 21)
    init = rows_as_initial_centroids(x, M, k, 22)[0]
    lo, hi = shard_rows(n, rank, world)
    cen = torch.from_numpy(init.copy())
    loss = kmeans_data_parallel(torch.from_numpy(x[lo:hi]), n, cen, iters, local_step=_np_local_step(o),
                                finalize=_np_finalize)
    if rank == 0:
        q.put((cen.numpy(), loss.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_kmeans_world2_matches_single_process(oracle):
    n, M, k, dsub, iters = 600, 3, 8, 4, 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, M, k, dsub, iters, q)) for r in range(2)]
    for p in procs:
        p.start()
    cen, loss = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: the oracle's sequential k-means from the same initial centroids
    x = normal((n, M * dsub), 21)
    init = rows_as_initial_centroids(x, M, k, 22)
    want_q, want_loss = oracle.train_pq(x, M, 3, iters, 1, init)
    assert np.linalg.norm(cen - want_q) / np.linalg.norm(want_q) < 1e-4
    assert np.allclose(loss, want_loss, rtol=1e-3)


# ---- chained mode: the running sums travel rank to rank, result bit-identical to one sequential pass -------------
def _np_assign(o):
    def assign(x_local, centroids):
        M, k, dsub = centroids.shape
        x, c = x_local.numpy(), centroids.numpy()
        return [o.cluster_assignments(c[m], np.ascontiguousarray(x[:, m * dsub:(m + 1) * dsub])).astype(np.int64)
                for m in range(M)]
    return assign


def _np_accumulate(x_local, centroids, codes, before, packed, m0=0, m1=None):
    """kmeans.rs:185-189 restated: one rounded f32 add per row, in row order, continuing from `before`, for the
    subquantizers [m0, m1) (packed layouts hold that slice only)."""
    M, k, dsub = centroids.shape
    m1 = M if m1 is None else m1
    ms = m1 - m0
    x = x_local.numpy()
    n0 = ms * k * dsub
    flat = np.zeros((n0 + ms * k + ms,), np.float32) if before is None else before.numpy().copy()
    sums, counts, sq = flat[:n0].reshape(ms, k, dsub), flat[n0:n0 + ms * k].reshape(ms, k), flat[n0 + ms * k:]
    for m in range(m0, m1):
        sub = x[:, m * dsub:(m + 1) * dsub]
        for i, a in enumerate(codes[m]):
            sums[m - m0, a] = sums[m - m0, a] + sub[i]  # float32 + float32, rounded once
            counts[m - m0, a] += np.float32(1)
        sq[m - m0] = np.float32(np.float64(sq[m - m0]) + (sub.astype(np.float64) ** 2).sum())
    packed.copy_(torch.from_numpy(flat))


def _np_finalize_f32(packed, n_total, centroids, loss):
    """kmeans.rs:191-197 in f32: centroid /= count for non-empty clusters, empty ones stay zero."""
    M, k, dsub = centroids.shape
    p = packed.numpy()
    n0 = M * k * dsub
    sums, counts = p[:n0].reshape(M, k, dsub), p[n0:n0 + M * k].reshape(M, k)
    c = np.zeros_like(sums)
    nz = counts > 0
    c[nz] = sums[nz] / counts[nz][:, None]
    centroids.copy_(torch.from_numpy(c))


def _chained_worker(rank, world, port, n, M, k, dsub, iters, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from reductive_b200.dist import kmeans_data_parallel, shard_rows

    o = orc.get()
    x = normal((n, M * dsub), 31)
    init = rows_as_initial_centroids(x, M, k, 32)[0]
    lo, hi = shard_rows(n, rank, world)
    cen = torch.from_numpy(init.copy())
    kmeans_data_parallel(torch.from_numpy(x[lo:hi]), n, cen, iters, mode="chained", assign=_np_assign(o),
                         accumulate=_np_accumulate, finalize=_np_finalize_f32)
    q.put((rank, cen.numpy()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_chained_data_parallel_kmeans_is_bit_identical_to_one_process(oracle, world):
    n, M, k, dsub, iters = 500, 5, 8, 4, 5
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_chained_worker, args=(r, world, port, n, M, k, dsub, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = normal((n, M * dsub), 31)
    init = rows_as_initial_centroids(x, M, k, 32)
    want_q, _ = oracle.train_pq(x, M, 3, iters, 1, init)
    for r in range(world):  # replicated and bit-identical to the oracle's sequential k-means
        assert np.array_equal(got[r].view(np.uint32), want_q.view(np.uint32)), r


# ---- the sharded mode's protocol (csrc/dist.cu) modelled over gloo ------------------------------------------
def _sharded_worker(rank, world, port, n, M, k, dsub, iters, q):
    """Same exchanges as rb_kmeans_dist_create / _iterate, with the oracle's CPU routines as the per-rank compute:
    column slices of x once, per iteration the assignments (to the owner of each subquantizer) and the centroids."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from reductive_b200.dist import shard_rows, subquantizer_range

    o = orc.get()
    x = normal((n, M * dsub), 21)
    cen = rows_as_initial_centroids(x, M, k, 22)[0].copy()
    lo, hi = shard_rows(n, rank, world)
    x_local = np.ascontiguousarray(x[lo:hi])
    ranges = [subquantizer_range(M, r, world) for r in range(world)]
    m0, m1 = ranges[rank]
    # create: everyone's rows of my columns, in rank (= row) order
    gathered = [None] * world
    dist.all_gather_object(gathered, [np.ascontiguousarray(x_local[:, a * dsub:b * dsub]) for a, b in ranges])
    xcol = np.concatenate([g[rank] for g in gathered], axis=0) if m1 > m0 else np.zeros((n, 0), np.float32)
    for _ in range(iters):
        codes_local = np.stack([o.cluster_assignments(cen[m], np.ascontiguousarray(x_local[:, m * dsub:(m + 1) * dsub]))
                                for m in range(M)]).astype(np.uint8)          # [M, n_local]: assignment by rows
        dist.all_gather_object(gathered, [codes_local[a:b] for a, b in ranges])
        codes_own = np.concatenate([g[rank] for g in gathered], axis=1)       # [m_own, n]: all rows, row order
        for j, m in enumerate(range(m0, m1)):                                  # ordered update by subquantizers
            cen[m] = o.update_centroids(cen[m], np.ascontiguousarray(xcol[:, j * dsub:(j + 1) * dsub]),
                                        codes_own[j].astype(np.uint64))
        dist.all_gather_object(gathered, cen[m0:m1])
        cen = np.concatenate([g for g in gathered if len(g)], axis=0)
    if rank == 0:
        q.put(cen)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,M", [(2, 3), (3, 4), (3, 2)])
def test_sharded_kmeans_protocol_is_bit_identical_to_one_process(oracle, world, M):
    """Assignment sharded by rows + update sharded by subquantizers reproduces the sequential run bit for bit
    (src/kmeans.rs:185-189 adds every cluster's rows in row order), including uneven row blocks and ranks that own no
    subquantizer."""
    n, k, dsub, iters = 601, 8, 4, 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, n, M, k, dsub, iters, q)) for r in range(world)]
    for p in procs:
        p.start()
    cen = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
    x = normal((n, M * dsub), 21)
    init = rows_as_initial_centroids(x, M, k, 22)
    want, _ = oracle.train_pq(x, M, 3, iters, 1, init, n_threads=2)
    assert np.array_equal(cen.view(np.int32), want.view(np.int32))
