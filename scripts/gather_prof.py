"""ncu target: one reconstruct_batch per shape (C2, C5, k = 16, C3)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import reductive_b200 as rb
for n, M, k, dsub in [(2_000_000, 30, 256, 10), (4_000_000, 16, 256, 8), (2_000_000, 16, 16, 8), (1_000_000, 96, 256, 8)]:
    q = np.random.default_rng(1).normal(size=(M, k, dsub)).astype(np.float32)
    pq = rb.Pq(None, q)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    codes = torch.randint(0, k, (n, M), generator=g, device="cuda", dtype=torch.uint8)
    rec = torch.empty((n, M * dsub), device="cuda")
    pq.reconstruct_batch_into(codes, rec)
    torch.cuda.synchronize()
    del rec, codes
print("done")
