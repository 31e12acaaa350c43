"""Data-parallel k-means on N GPUs (torchrun): all-reduce mode vs chained mode vs a one-GPU run of the same data.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      scripts/dist_check.py [rows] [iters]
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402
from reductive_b200.dist import kmeans_data_parallel, shard_rows  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
M, k, dsub = 96, 256, 8
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
g = torch.Generator(device="cuda"); g.manual_seed(77)          # same data on every rank, each keeps its shard
x = torch.randn((n, M * dsub), generator=g, device="cuda")
c0 = x[torch.randperm(n, generator=g, device="cuda")[:k]].reshape(k, M, dsub).permute(1, 0, 2).contiguous()
lo, hi = shard_rows(n, rank, world)
xl = x[lo:hi].contiguous()
res = {}
for mode in ("allreduce", "chained"):
    cen = c0.clone()
    kmeans_data_parallel(xl, n, cen, 1, mode=mode)
    cen = c0.clone()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kmeans_data_parallel(xl, n, cen, iters, mode=mode)
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[mode] = (cen, t.item())
if rank == 0:
    ref = c0.clone()
    os.environ.pop("WORLD_SIZE", None)
    pl = torch.empty((M * k * dsub + M * k + M,), device="cuda")
    from reductive_b200.dist import cuda_finalize, cuda_local_step
    loss = torch.zeros((M,), device="cuda")
    for _ in range(iters):
        cuda_local_step(x, ref, pl)
        cuda_finalize(pl, n, ref, loss)
    for mode, (cen, ms) in res.items():
        rel = ((cen - ref).norm() / ref.norm()).item()
        same = bool(torch.equal(cen.view(torch.int32), ref.view(torch.int32)))
        print(f"{world} GPUs, {mode}: {ms:.3f} ms/iter; vs one-GPU run after {iters} iterations: rel {rel:.3e}, bit-identical {same}",
              flush=True)
dist.barrier()
dist.destroy_process_group()
