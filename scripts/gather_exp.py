"""Experiments on the gather kernel: alignment / lookup-shape sensitivity."""
import sys, os
import numpy as np, torch
sys.path.insert(0, ".")
import reductive_b200 as rb
def run(name, n, M, k, dsub, ldo=None):
    d = M * dsub
    ldo = ldo or d
    q = np.random.default_rng(1).normal(size=(M, k, dsub)).astype(np.float32)
    pq = rb.Pq(None, q)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    codes = torch.randint(0, k, (n, M), generator=g, device="cuda", dtype=torch.uint8)
    buf = torch.empty((n, ldo), device="cuda")
    rec = buf[:, :d]
    pq.reconstruct_batch_into(codes, rec); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): pq.reconstruct_batch_into(codes, rec)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = n * (M + 4 * d) / ms / 1e6
    print(f"{name}: n={n} M={M} k={k} dsub={dsub} ldo={ldo}: {ms:.3f} ms, {gbs:.0f} GB/s ({gbs/6543.4:.2f})", flush=True)
run("1 group, row 400B", 6_000_000, 10, 256, 10)
run("1 group, row 256B", 8_000_000, 8, 256, 8)
run("1 group, row 384B", 6_000_000, 12, 256, 8)
run("1 group, row 360B d6", 6_000_000, 15, 256, 6)
run("2 groups, row 800B", 3_000_000, 20, 256, 10)
run("C2", 2_000_000, 30, 256, 10)
