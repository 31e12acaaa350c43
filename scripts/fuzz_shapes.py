"""Random shapes through every fast path against the exact kernels on the same GPU (no CPU oracle: sizes are large):
tensor encode vs exact encode (codes equal), tiled gather vs index_select (bits equal), tensor rotation decode vs the
exact GEMM (<= 1e-5 of max), tensor rotation + tensor encode vs exact rotation + exact encode (codes equal).

  python scripts/fuzz_shapes.py [n_configs] [seed]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
DSUBS = [2, 4, 6, 8, 10, 12, 16, 20, 24, 30, 32]
bad = 0
t0 = time.time()
for it in range(n_cfg):
    dsub = int(rng.choice(DSUBS))
    M = int(rng.integers(1, max(2, min(96, 1024 // dsub)) + 1))
    if (M * dsub) % 4:
        M += 1 if (dsub % 4 == 2) else 0
    if (M * dsub) % 4:
        continue
    k = int(rng.choice([256, 256, 200, 128, 65]))
    n = int(rng.choice([1_500, 20_000, 150_000, 400_000]))
    n = n + int(rng.integers(0, 257))
    d = M * dsub
    if n * d > 250_000_000:
        n = 250_000_000 // d
    q = rng.normal(size=(M, k, dsub)).astype(np.float32) * float(rng.choice([1e-3, 1.0, 50.0]))
    g = torch.Generator(device="cuda")
    g.manual_seed(int(rng.integers(1 << 30)))
    x = torch.randn((n, d), generator=g, device="cuda") * float(np.abs(q).mean() * 2)
    msgs = []
    # plain quantizer
    pq = rb.Pq(None, q)
    ce = torch.empty((n, M), dtype=torch.uint8, device="cuda")
    ct = torch.empty((n, M), dtype=torch.uint8, device="cuda")
    rb.set_encode_algo(rb.ENCODE_EXACT)
    pq.quantize_batch_into(x, ce)
    rb.set_encode_algo(rb.ENCODE_AUTO)
    pq.quantize_batch_into(x, ct)
    diff = int((ce != ct).sum())
    if diff:
        msgs.append(f"encode: {diff} codes differ")
    rec = torch.empty((n, d), device="cuda")
    pq.reconstruct_batch_into(ce, rec)
    qd = torch.from_numpy(q).cuda()
    for m in range(0, M, max(1, M // 3)):
        if not torch.equal(rec[:, m * dsub:(m + 1) * dsub], qd[m].index_select(0, ce[:, m].long())):
            msgs.append(f"gather: subquantizer {m} differs")
            break
    # projected quantizer
    if d >= 32 and d <= 640:
        r = np.linalg.qr(rng.normal(size=(d, d)))[0].astype(np.float32)
        pp = rb.Pq(r, q)
        rb.set_project_algo(rb.PROJECT_EXACT)
        rb.set_encode_algo(rb.ENCODE_EXACT)
        pp.quantize_batch_into(x, ce)
        re_ = torch.empty((n, d), device="cuda")
        pp.reconstruct_batch_into(ce, re_)
        rb.set_project_algo(rb.PROJECT_AUTO)
        rb.set_encode_algo(rb.ENCODE_AUTO)
        pp.quantize_batch_into(x, ct)
        pp.reconstruct_batch_into(ce, rec)
        diff = int((ce != ct).sum())
        if diff:
            msgs.append(f"projected encode: {diff} codes differ")
        err = float((rec - re_).abs().max() / re_.abs().max())
        if not err <= 1e-5:
            msgs.append(f"projected decode: error {err:.2e} of max")
        del re_
    torch.cuda.synchronize()
    status = "ok" if not msgs else "FAIL " + "; ".join(msgs)
    bad += bool(msgs)
    print(f"[{it:3d}] n={n} M={M} k={k} dsub={dsub} d={d}: {status}", flush=True)
    del x, ce, ct, rec
    if time.time() - t0 > 240:
        break
print(f"{bad} failing configurations")
sys.exit(1 if bad else 0)
