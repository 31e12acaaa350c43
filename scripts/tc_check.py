"""Tensor encode kernel vs exact kernel on the BASELINE shapes: equality of all codes, flagged-pair rate, timings.

  python scripts/tc_check.py [rows]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
for name, n, M, dsub in [("C2", rows, 30, 10), ("C3", rows // 2, 96, 8), ("C5", rows * 2, 16, 8), ("C1", 10_000, 10, 30),
                         ("d16", rows // 2, 8, 16), ("d20", rows // 2, 15, 20), ("d12", rows // 2, 25, 12), ("d4", rows, 32, 4)]:
    k, d = 256, M * dsub
    g = torch.Generator(device="cuda")
    g.manual_seed(1000)
    x = torch.randn((n, d), generator=g, device="cuda")
    q = np.random.default_rng(1).normal(size=(M, k, dsub)).astype(np.float32)
    pq = rb.Pq(None, q)
    out = {}
    for algo, nm in [(rb.ENCODE_EXACT, "exact"), (rb.ENCODE_TENSOR, "tensor")]:
        rb.set_encode_algo(algo)
        codes = torch.empty((n, M), dtype=torch.uint8, device="cuda")
        pq.quantize_batch_into(x, codes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            pq.quantize_batch_into(x, codes)
        e1.record()
        torch.cuda.synchronize()
        out[nm] = (codes, e0.elapsed_time(e1) / 3)
    diff = int((out["exact"][0] != out["tensor"][0]).sum().item())
    print(f"{name}: n={n} M={M} dsub={dsub}: exact {out['exact'][1]:.3f} ms, tensor {out['tensor'][1]:.3f} ms, "
          f"{diff} of {n * M} codes differ", flush=True)
    del x
rb.set_encode_algo(rb.ENCODE_AUTO)
