import sys
import numpy as np, torch
sys.path.insert(0, ".")
import reductive_b200 as rb
n, d, M = 1_000_000, 300, 30
g = torch.Generator(device="cuda"); g.manual_seed(5)
x = torch.randn((n, d), generator=g, device="cuda")
q = np.random.default_rng(1).normal(size=(M, 256, d // M)).astype(np.float32)
R = np.linalg.qr(np.random.default_rng(2).normal(size=(d, d)))[0].astype(np.float32)
pq = rb.Pq(np.ascontiguousarray(R), q)
codes = torch.empty((n, M), dtype=torch.uint8, device="cuda"); rec = torch.empty((n, d), device="cuda")
def timed(f):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
print(f"C4 encode {timed(lambda: pq.quantize_batch_into(x, codes)):.3f} ms, decode {timed(lambda: pq.reconstruct_batch_into(codes, rec)):.3f} ms")
