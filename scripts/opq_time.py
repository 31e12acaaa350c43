"""Times the device parts of OPQ training on the C4 shape (1 M x 300, M = 30, k = 256): rb_covariance (PCA init,
linalg.rs:23-44) and rb_opq_train_iteration (opq.rs:161-189 without the d x d SVD)."""
import sys

import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402,F401
from reductive_b200._cabi import check, lib  # noqa: E402

n, d, M, k = 1_000_000, 300, 30, 256
if len(sys.argv) > 1:
    n = int(sys.argv[1])
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((n, d), generator=g, device="cuda")
proj = torch.linalg.qr(torch.randn((d, d), generator=g, device="cuda"))[0].contiguous()
cen = x[torch.randperm(n, device="cuda", generator=g)[: M * k]].reshape(k, M, d // M, M)[:, 0].permute(1, 0, 2)[:M].contiguous()
cen = torch.randn((M, k, d // M), generator=g, device="cuda")
cov = torch.empty((d, d), device="cuda")
xty = torch.empty((d, d), device="cuda")
st = torch.cuda.current_stream().cuda_stream


def timed(f, reps=5):
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tc = timed(lambda: check(lib.rb_covariance(x.data_ptr(), n, d, x.stride(0), cov.data_ptr(), st)))
c2 = cen.clone()
ti = timed(lambda: check(lib.rb_opq_train_iteration(x.data_ptr(), n, d, x.stride(0), proj.data_ptr(), c2.data_ptr(), M, k,
                                                    xty.data_ptr(), st)))
ref = ((x - x.mean(0)).double().T @ (x - x.mean(0)).double() / (n - 1)).float()
print(f"n={n} d={d}: covariance {tc:.3f} ms ({2 * n * d * d / tc / 1e9:.1f} TFLOP/s, max err {float((cov - ref).abs().max()):.2e}); "
      f"opq train iteration (device part) {ti:.3f} ms", flush=True)
