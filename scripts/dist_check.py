"""Data-parallel k-means on N GPUs (torchrun): the sharded mode (rows for assignment, subquantizers for the ordered
update; NCCL from the C++ library) against a one-GPU run of the same data.  ASSERTS bit-identity.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/dist_check.py [rows] [iters] [--also-legacy]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402,F401
from reductive_b200.dist import (Comm, ShardedKMeans, cuda_finalize, cuda_local_step, kmeans_data_parallel,  # noqa: E402
                                 shard_rows)

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 1_000_000
iters = int(args[1]) if len(args) > 1 else 10
legacy = "--also-legacy" in sys.argv
M, k, dsub = 96, 256, 8
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def block(r):
    """rows of rank r (the same on whichever device generates them)"""
    lo, hi = shard_rows(n, r, world)
    g = torch.Generator(device="cuda")
    g.manual_seed(77 + r)
    return torch.randn((hi - lo, M * dsub), generator=g, device="cuda")


xl = block(rank)
# initial centroids = rows (7919 j + 104729 m) mod n; here taken from a fixed draw so every rank agrees
gi = torch.Generator(device="cuda")
gi.manual_seed(5)
c0 = torch.randn((M, k, dsub), generator=gi, device="cuda")

comm = Comm()
km = ShardedKMeans(comm, xl, M, k, dsub)
cen = c0.clone()
km.iterate(cen)  # warm-up
cen = c0.clone()
loss = torch.zeros((M,), device="cuda")
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    km.iterate(cen, loss)
e1.record()
dist.barrier()
torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
res = {"sharded": (cen, t.item())}
if rank == 0:
    print(f"code exchange: {'peer-mapped window (stores into the owners\' memory)' if km.peer_window else 'ncclSend/Recv'}", flush=True)
km.close()

if legacy:
    for mode in ("allreduce", "chained"):
        c = c0.clone()
        kmeans_data_parallel(xl, n, c, 1, mode=mode)
        c = c0.clone()
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        kmeans_data_parallel(xl, n, c, iters, mode=mode)
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = (c, t.item())

# every rank must hold the same centroids
ref_bits = cen.view(torch.int32).clone()
dist.broadcast(ref_bits, src=0)
assert torch.equal(ref_bits, cen.view(torch.int32)), f"rank {rank}: centroids differ from rank 0's"

if rank == 0:
    x = torch.cat([block(r) for r in range(world)])  # one-GPU run of the same rows
    ref = c0.clone()
    pl = torch.empty((M * k * dsub + M * k + M,), device="cuda")
    l1 = torch.zeros((M,), device="cuda")
    for _ in range(iters):
        cuda_local_step(x, ref, pl)
        cuda_finalize(pl, n, ref, l1)
    for mode, (c, ms) in res.items():
        rel = ((c - ref).norm() / ref.norm()).item()
        same = bool(torch.equal(c.view(torch.int32), ref.view(torch.int32)))
        print(f"{world} GPUs, {mode}: {ms:.3f} ms/iter; vs one-GPU run after {iters} iterations: rel {rel:.3e}, "
              f"bit-identical {same}", flush=True)
        if mode in ("sharded", "chained"):
            assert same, f"{mode} mode is not bit-identical to the one-GPU run"
    assert torch.allclose(loss, l1, rtol=1e-4), "loss differs from the one-GPU run"
    print("dist_check ok", flush=True)
dist.barrier()
comm.close()
dist.destroy_process_group()
