"""Short ncu target: the C2 encode (2M x 300, 30 x 256) through the tensor kernel, three calls."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
M, k, dsub = 30, 256, 10
rb.set_encode_algo(rb.ENCODE_TENSOR)
q = np.random.default_rng(1).normal(size=(M, k, dsub)).astype(np.float32)
pq = rb.Pq(None, q)
g = torch.Generator(device="cuda")
g.manual_seed(1000)
x = torch.randn((rows, M * dsub), generator=g, device="cuda")
codes = torch.empty((rows, M), dtype=torch.uint8, device="cuda")
for _ in range(3):
    pq.quantize_batch_into(x, codes)
torch.cuda.synchronize()
print("tc profile target done", rb.kernel_launch_count(), "launches")
