"""Tensor-core rotation vs the exact FP32 GEMM: decode (reconstruct_batch of a projected quantizer) error and time,
encode codes equality and time.

  python scripts/proj_tc_check.py [rows] [C4,d768,...]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402
from reductive_b200 import _cabi  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None


def timed(f, reps=3):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, n, M, k, dsub in [("C4", rows, 30, 256, 10), ("d768", rows // 4, 96, 256, 8), ("d128", rows, 16, 256, 8),
                            ("d36", 50_000, 4, 128, 9)]:
    if only and name not in only:
        continue
    d = M * dsub
    rng = np.random.default_rng(5)
    q = rng.normal(size=(M, k, dsub)).astype(np.float32)
    r = np.linalg.qr(rng.normal(size=(d, d)))[0].astype(np.float32)
    pq = rb.Pq(r, q)
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    codes = torch.randint(0, k, (n, M), generator=g, device="cuda", dtype=torch.uint8)
    rec_e = torch.empty((n, d), device="cuda")
    rec_t = torch.empty((n, d), device="cuda")
    _cabi.set_project_algo(_cabi.PROJECT_EXACT)
    ms_e = timed(lambda: pq.reconstruct_batch_into(codes, rec_e))
    _cabi.set_project_algo(_cabi.PROJECT_AUTO)
    ms_t = timed(lambda: pq.reconstruct_batch_into(codes, rec_t))
    err = float((rec_e - rec_t).abs().max() / rec_e.abs().max())
    if err > 1e-4:
        e = (rec_e - rec_t).abs()
        print("   err by column block of 32:", [f"{float(e[:, c:c + 32].max()):.1e}" for c in range(0, d, 32)])
        print("   err by row quarter of the first tile:", [f"{float(e[r:r + 32].max()):.1e}" for r in range(0, 128, 32)])
        print("   err by tile (first 6):", [f"{float(e[t * 128:(t + 1) * 128].max()):.1e}" for t in range(6)])
        bad_rows = (e.max(dim=1).values > 1e-3).float().mean()
        print(f"   fraction of rows with an error: {float(bad_rows):.4f}")
    x = torch.randn((n, d), generator=g, device="cuda")
    ce = torch.empty((n, M), device="cuda", dtype=torch.uint8)
    ct = torch.empty((n, M), device="cuda", dtype=torch.uint8)
    _cabi.set_project_algo(_cabi.PROJECT_EXACT)
    ms_ee = timed(lambda: pq.quantize_batch_into(x, ce))
    _cabi.set_project_algo(_cabi.PROJECT_AUTO)
    ms_et = timed(lambda: pq.quantize_batch_into(x, ct))
    diff = int((ce != ct).sum())
    print(f"{name}: n={n} d={d}: decode exact {ms_e:.3f} ms, tensor {ms_t:.3f} ms, max err {err:.2e} of max | "
          f"encode exact {ms_ee:.3f} ms, tensor {ms_et:.3f} ms, {diff} codes differ", flush=True)
    del rec_e, rec_t, codes, x
_cabi.set_project_algo(_cabi.PROJECT_AUTO)
