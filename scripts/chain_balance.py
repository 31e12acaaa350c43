"""How the ordered centroid update (sort + chains) depends on the cluster-size distribution: C3 shape, codes forced to
(a) perfectly balanced, (b) the k-means assignment of Gaussian data, (c) one cluster holding 10 % of every chunk."""
import sys
import torch
sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: F401
from reductive_b200.dist import cuda_accumulate, cuda_assign
from reductive_b200._cabi import lib
M, k, dsub, n = 96, 256, 8, 1_000_000
g = torch.Generator(device="cuda"); g.manual_seed(77)
x = torch.randn((n, M * dsub), generator=g, device="cuda")
cen = torch.randn((M, k, dsub), generator=g, device="cuda")
packed = torch.empty((M * k * dsub + M * k + M,), device="cuda")
codes = cuda_assign(x, cen)
pitch = int(lib.rb_kmeans_code_pitch(n))
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cv = codes.view(M, pitch)
sizes = torch.bincount(cv[0, :n].long(), minlength=k).float()
print(f"k-means assignment: largest cluster {float(sizes.max() / sizes.mean()):.1f}x the mean: "
      f"accumulate {t(lambda: cuda_accumulate(x, cen, codes, None, packed)):.3f} ms")
bal = codes.clone(); bv = bal.view(M, pitch)
bv[:, :n] = (torch.arange(n, device="cuda") % k).to(torch.uint8)[None, :]
print(f"balanced: accumulate {t(lambda: cuda_accumulate(x, cen, bal, None, packed)):.3f} ms")
sk = bal.clone(); sv = sk.view(M, pitch)
mask = (torch.arange(n, device="cuda") % 10) == 0
sv[:, :n][:, mask] = 0
print(f"one cluster with 10 % of the rows: accumulate {t(lambda: cuda_accumulate(x, cen, sk, None, packed)):.3f} ms")
