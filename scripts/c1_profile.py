"""C1: Pq train (10 iters, 1 attempt) on 10k x 300, M = 10 — launch-level timeline target."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import reductive_b200 as rb
x = np.random.default_rng(0).normal(size=(10_000, 300)).astype(np.float32)
rng = np.random.default_rng(3)
for _ in range(3):
    t0 = time.perf_counter()
    pq = rb.Pq.train_pq_using(10, 8, 10, 1, x, rng)
    print(f"train {1e3 * (time.perf_counter() - t0):.2f} ms")
