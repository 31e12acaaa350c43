"""Split timing of one C3-shaped k-means iteration: assign | accumulate (sort + chains) | finalize, at several n."""
import sys
import torch
sys.path.insert(0, ".")
import reductive_b200 as rb
from reductive_b200.dist import cuda_accumulate, cuda_assign, cuda_finalize
M, k, dsub = 96, 256, 8
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for n in (1_000_000, 262_144, 32_768):
    g = torch.Generator(device="cuda"); g.manual_seed(77)
    x = torch.randn((n, M * dsub), generator=g, device="cuda")
    cen = torch.randn((M, k, dsub), generator=g, device="cuda")
    packed = torch.empty((M * k * dsub + M * k + M,), device="cuda")
    loss = torch.zeros((M,), device="cuda")
    codes = cuda_assign(x, cen)
    print(f"n={n}: assign {t(lambda: cuda_assign(x, cen)):.3f} ms | accumulate {t(lambda: cuda_accumulate(x, cen, codes, None, packed)):.3f} ms | "
          f"finalize {t(lambda: cuda_finalize(packed, n, cen.clone(), loss)):.3f} ms", flush=True)
