"""A/B timing helper: C2/C3/C5 tensor encode time (ms) with the library selected by RB_LIB_PATH."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import reductive_b200 as rb
rb.set_encode_algo(rb.ENCODE_TENSOR)
out = []
for name, n, M, dsub in [("C2", 2_000_000, 30, 10), ("C3", 1_000_000, 96, 8), ("C5", 4_000_000, 16, 8)]:
    g = torch.Generator(device="cuda"); g.manual_seed(1000)
    x = torch.randn((n, M * dsub), generator=g, device="cuda")
    pq = rb.Pq(None, np.random.default_rng(1).normal(size=(M, 256, dsub)).astype(np.float32))
    codes = torch.empty((n, M), dtype=torch.uint8, device="cuda")
    for _ in range(3): pq.quantize_batch_into(x, codes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): pq.quantize_batch_into(x, codes)
    e1.record(); torch.cuda.synchronize()
    out.append(f"{name} {e0.elapsed_time(e1) / 10:.3f}")
    del x, codes
print(" | ".join(out), flush=True)
