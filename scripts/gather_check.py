"""reconstruct_batch on the BASELINE shapes: equality with an index_select reference on the device, and timing.

  python scripts/gather_check.py [rows]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
PEAK = 6543.4
for name, n, M, k, dsub in [("C2", rows, 30, 256, 10), ("C3", rows // 2, 96, 256, 8), ("C5", rows * 2, 16, 256, 8),
                            ("C1", 10_000, 10, 256, 30), ("k16", rows, 16, 16, 8), ("d6", rows // 2, 50, 200, 6),
                            ("ragged", 1_000_003, 30, 256, 10)]:
    d = M * dsub
    q = np.random.default_rng(1).normal(size=(M, k, dsub)).astype(np.float32)
    pq = rb.Pq(None, q)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    codes = torch.randint(0, k, (n, M), generator=g, device="cuda", dtype=torch.uint8)
    rec = torch.empty((n, d), device="cuda")
    pq.reconstruct_batch_into(codes, rec)
    torch.cuda.synchronize()
    qd = torch.from_numpy(q).cuda()
    ok = True
    for m in range(M):
        want = qd[m].index_select(0, codes[:, m].long())
        ok &= bool(torch.equal(rec[:, m * dsub:(m + 1) * dsub], want))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pq.reconstruct_batch_into(codes, rec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = n * (M + 4 * d) / ms / 1e6
    print(f"{name}: n={n} M={M} k={k} dsub={dsub}: {ms:.3f} ms, {gbs:.0f} GB/s ({gbs / PEAK:.2f} of copy peak), equal={ok}",
          flush=True)
    del rec, codes
