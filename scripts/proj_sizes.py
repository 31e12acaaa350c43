import sys, numpy as np, torch
sys.path.insert(0, ".")
import reductive_b200 as rb
from reductive_b200 import _cabi
def run(n, M, k, dsub):
    d = M * dsub
    rng = np.random.default_rng(5)
    q = rng.normal(size=(M, k, dsub)).astype(np.float32)
    r = np.linalg.qr(rng.normal(size=(d, d)))[0].astype(np.float32)
    pq = rb.Pq(r, q)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    codes = torch.randint(0, k, (n, M), generator=g, device="cuda", dtype=torch.uint8)
    rec = torch.empty((n, d), device="cuda")
    try:
        pq.reconstruct_batch_into(codes, rec); torch.cuda.synchronize()
        print(f"n={n} d={d}: ok", flush=True)
    except Exception as e:
        print(f"n={n} d={d}: FAIL {str(e)[:80]}", flush=True)
        sys.exit(1)
for n in (2000, 20000, 100000, 300000, 1000000):
    run(n, 16, 256, 8)
