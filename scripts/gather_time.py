"""One line of reconstruct_batch timings (ms) on C2 / C3 / C5 / d6 / k16, each checked against index_select.

  [RB_GATHER_BIG=0|1] [RB_GATHER_W=..] [RB_GATHER_STAGES=..] python scripts/gather_time.py
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402

parts = []
for name, n, M, k, dsub in [("C2", 2_000_000, 30, 256, 10), ("C3", 1_000_000, 96, 256, 8), ("C5", 4_000_000, 16, 256, 8),
                            ("d6", 1_000_000, 50, 200, 6), ("k16", 2_000_000, 16, 16, 8)]:
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
    ok = all(bool(torch.equal(rec[:, m * dsub:(m + 1) * dsub], qd[m].index_select(0, codes[:, m].long())))
             for m in range(0, M, 7))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        pq.reconstruct_batch_into(codes, rec)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    parts.append(f"{name} {ms:.3f} ({n * (M + 4 * d) / ms / 1e6 / 6543.4:.2f}){'' if ok else ' WRONG'}")
    del rec, codes
print(" | ".join(parts), flush=True)
