"""Times the fused decode + dot (rb_qstore_dot) on the C2 store: 2 M rows x 30 u8 codes, nq queries of 300 floats,
against the unfused route (reconstruct_batch into a [n, d] matrix, then a library GEMM)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402

n, M, k, dsub = 2_000_000, 30, 256, 10
if len(sys.argv) > 2:
    n, M, dsub = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = M * dsub
g = torch.Generator(device="cuda").manual_seed(1)
q = torch.randn((M, k, dsub), generator=g, device="cuda")
codes = torch.randint(0, k, (n, M), dtype=torch.uint8, device="cuda", generator=g)
norms = torch.rand((n,), generator=g, device="cuda") + 0.5
pq = rb.Pq(None, q.cpu().numpy())
store = rb.QuantizedArray(pq, codes, norms)


def timed(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for nq in (1, 8, 64, 256):
    if nq * n * 4 > 8e9:
        continue
    queries = torch.randn((nq, d), generator=g, device="cuda")
    out = torch.empty((nq, n), device="cuda")
    t = timed(lambda: store.dot(queries, out))
    rec = torch.empty((n, d), device="cuda")

    def unfused():
        pq.reconstruct_batch_into(codes, rec)
        torch.matmul(queries, (rec * norms[:, None]).T, out=out)

    tu = timed(unfused, 3)
    passes = -(-nq // 8)
    print(f"nq={nq}: fused {t:.3f} ms = {nq * n / t / 1e6:.1f} G scores/s, {nq * n * M / t / 1e6:.0f} G lookups/s; "
          f"HBM bytes {(passes * n * M + nq * n * 4) / 1e9:.2f} GB -> {(passes * n * M + nq * n * 4) / t / 1e6:.0f} GB/s; "
          f"unfused {tu:.3f} ms", flush=True)
