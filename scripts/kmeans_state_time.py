"""One-GPU k-means iteration through the library's training state (streaming update) vs the stateless entry points."""
import sys

import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402
from reductive_b200.dist import Comm, ShardedKMeans, cuda_finalize, cuda_local_step  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M, k, dsub = (int(sys.argv[2]), 256, int(sys.argv[3])) if len(sys.argv) > 3 else (96, 256, 8)
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.randn((n, M * dsub), generator=g, device="cuda")
c0 = x[torch.randperm(n, generator=g, device="cuda")[:k]].reshape(k, M, dsub).permute(1, 0, 2).contiguous()
iters = 10
comm = Comm(rank=0, world=1)
res = {}
for mode in (3, 1):
    rb.set_kmeans_update(mode)
    km = ShardedKMeans(comm, x, M, k, dsub)
    cen = c0.clone()
    km.iterate(cen)
    cen = c0.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        km.iterate(cen)
    e1.record()
    torch.cuda.synchronize()
    res[mode] = (cen, e0.elapsed_time(e1) / iters)
    km.close()
rb.set_kmeans_update(1)
ref = c0.clone()
pl = torch.empty((M * k * dsub + M * k + M,), device="cuda")
for _ in range(iters):
    cuda_local_step(x, ref, pl)
    cuda_finalize(pl, n, ref, None)
for mode, (cen, ms) in res.items():
    print(f"n={n} M={M} dsub={dsub}: update mode {mode}: {ms:.3f} ms/iter, bit-identical to the stateless path: "
          f"{bool(torch.equal(cen.view(torch.int32), ref.view(torch.int32)))}", flush=True)
