"""BASELINE.json's five configurations on one GPU (C5: one of the eight 12.5M-row shards), device-resident timing.

  python scripts/config_sweep.py
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402
from reductive_b200.dist import kmeans_data_parallel  # noqa: E402


def timed(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


g = torch.Generator(device="cuda"); g.manual_seed(5)
PEAK = 6543.4
# C1: Pq train (10 iters, 1 attempt) + quantize_batch on 10k x 300, 10 subquantizers x 256 centroids
x1 = torch.randn((10_000, 300), generator=g, device="cuda")
x1h = x1.cpu().numpy()
rng = np.random.default_rng(3)
t0 = time.perf_counter()
pq1 = rb.Pq.train_pq_using(10, 8, 10, 1, x1h, rng)
t_train = time.perf_counter() - t0
t0 = time.perf_counter()
pq1 = rb.Pq.train_pq_using(10, 8, 10, 1, x1h, rng)
t_train = min(t_train, time.perf_counter() - t0)
c1 = torch.empty((10_000, 10), dtype=torch.uint8, device="cuda")
ms = timed(lambda: pq1.quantize_batch_into(x1, c1))
print(f"C1: train 10 iters (host array in, handle out) {t_train * 1e3:.1f} ms wall; quantize_batch 10k x 300: {ms:.3f} ms", flush=True)

# C2 is bench.py's headline.  C4: projected encode + decode, 1M x 300, M = 30 and M = 10
for M in (30, 10):
    n, d = 1_000_000, 300
    dsub = d // M
    x = torch.randn((n, d), generator=g, device="cuda")
    q = np.random.default_rng(1).normal(size=(M, 256, dsub)).astype(np.float32)
    R = np.linalg.qr(np.random.default_rng(2).normal(size=(d, d)))[0].astype(np.float32)
    pq = rb.Pq(np.ascontiguousarray(R), q)
    codes = torch.empty((n, M), dtype=torch.uint8, device="cuda")
    rec = torch.empty((n, d), device="cuda")
    me = timed(lambda: pq.quantize_batch_into(x, codes))
    md = timed(lambda: pq.reconstruct_batch_into(codes, rec))
    print(f"C4 (M={M}): projected encode {me:.3f} ms ({n / me / 1e3:.0f} Mvec/s), projected decode {md:.3f} ms; "
          f"projection alone is 2*n*d^2 = {2 * n * d * d / 1e9:.0f} GFLOP in exact FP32 order", flush=True)
    del x, rec, codes

# C3 one-GPU iteration
n, M, dsub = 1_000_000, 96, 8
x = torch.randn((n, M * dsub), generator=g, device="cuda")
cen = torch.randn((M, 256, dsub), generator=g, device="cuda")
ms = timed(lambda: kmeans_data_parallel(x, n, cen, 1))
print(f"C3: k-means iteration 1M x 768, 96 x 256: {ms:.3f} ms/iter (bit-exact ordered update)", flush=True)
del x

# C5: one 12.5M x 128 shard, M = 16
n, M, dsub = 12_500_000, 16, 8
x = torch.randn((n, M * dsub), generator=g, device="cuda")
pq = rb.Pq(None, np.random.default_rng(1).normal(size=(M, 256, dsub)).astype(np.float32))
codes = torch.empty((n, M), dtype=torch.uint8, device="cuda")
ms = timed(lambda: pq.quantize_batch_into(x, codes), 3)
gbs = n * (4 * M * dsub + M) / ms / 1e6
print(f"C5 shard: quantize_batch 12.5M x 128, 16 x 256: {ms:.3f} ms = {n / ms / 1e6:.2f} Gvec/s per GPU, {gbs:.0f} GB/s "
      f"({gbs / PEAK:.2f} of HBM copy peak)", flush=True)
rec = torch.empty((n, M * dsub), device="cuda")
ms = timed(lambda: pq.reconstruct_batch_into(codes, rec), 3)
gbs = n * (4 * M * dsub + M) / ms / 1e6
print(f"C5 shard: reconstruct_batch: {ms:.3f} ms, {gbs:.0f} GB/s ({gbs / PEAK:.2f} of HBM copy peak)", flush=True)
