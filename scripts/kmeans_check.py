"""C3-shaped k-means step timing: ordered (reference-order, bit-exact) vs atomic centroid update; centroid agreement."""
import sys
import torch
sys.path.insert(0, ".")
import reductive_b200 as rb
from reductive_b200.dist import kmeans_data_parallel
n, M, k, dsub = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 96, 256, 8
g = torch.Generator(device="cuda"); g.manual_seed(77)
x = torch.randn((n, M * dsub), generator=g, device="cuda")
c0 = x[torch.randperm(n, generator=g, device="cuda")[:k]].reshape(k, M, dsub).permute(1, 0, 2).contiguous()
res = {}
for name, ordered in (("ordered", True), ("atomic", False)):
    rb.set_kmeans_update(ordered)
    cen = c0.clone()
    kmeans_data_parallel(x, n, cen, 1)
    torch.cuda.synchronize()
    cen = c0.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    kmeans_data_parallel(x, n, cen, 10)
    e1.record(); torch.cuda.synchronize()
    res[name] = cen
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms/iter", flush=True)
rel = (res["ordered"] - res["atomic"]).norm() / res["ordered"].norm()
print(f"relative difference after 10 iterations (atomic vs ordered): {rel.item():.3e}")
rb.set_kmeans_update(True)
