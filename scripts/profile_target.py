"""Short program for ncu: one pass of each hot kernel at BASELINE config sizes (C2 encode/decode, C3 k-means step).

  python scripts/profile_target.py [exact|tensor|auto] [rows]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402
from reductive_b200.dist import kmeans_data_parallel  # noqa: E402

algo = sys.argv[1] if len(sys.argv) > 1 else "auto"
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
rb.set_encode_algo({"auto": rb.ENCODE_AUTO, "exact": rb.ENCODE_EXACT, "tensor": rb.ENCODE_TENSOR}[algo])
M, k, dsub = 30, 256, 10
q = np.random.default_rng(1).normal(size=(M, k, dsub)).astype(np.float32)
pq = rb.Pq(None, q)
g = torch.Generator(device="cuda")
g.manual_seed(1000)
x = torch.randn((rows, M * dsub), generator=g, device="cuda")
codes = torch.empty((rows, M), dtype=torch.uint8, device="cuda")
rec = torch.empty((rows, M * dsub), device="cuda")
for _ in range(2):
    pq.quantize_batch_into(x, codes)
    pq.reconstruct_batch_into(codes, rec)
torch.cuda.synchronize()
del rec
n3 = rows // 2
x3 = torch.randn((n3, 768), generator=g, device="cuda")
cen = torch.randn((96, 256, 8), generator=g, device="cuda")
kmeans_data_parallel(x3, n3, cen, 2)
torch.cuda.synchronize()
print("profile target done", rb.kernel_launch_count(), "launches")
