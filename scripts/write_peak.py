"""Write-only and read-only HBM ceilings on this box (context for the gather kernel's roofline fraction)."""
import torch
def t(f, n=10):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
N = 600_000_000  # 2.4 GB of f32
a = torch.empty(N, device="cuda"); b = torch.empty(N, device="cuda")
ms = t(lambda: a.fill_(1.0)); print(f"fill_ 2.4 GB: {ms:.3f} ms -> {4*N/ms/1e6:.0f} GB/s write-only")
ms = t(lambda: a.zero_()); print(f"zero_ (memset) 2.4 GB: {ms:.3f} ms -> {4*N/ms/1e6:.0f} GB/s write-only")
ms = t(lambda: b.copy_(a)); print(f"copy 2.4 GB: {ms:.3f} ms -> {8*N/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: a.sum()); print(f"sum 2.4 GB: {ms:.3f} ms -> {4*N/ms/1e6:.0f} GB/s read-only")
a2 = a.view(2_000_000, 300)
ms = t(lambda: a2[:, :100].fill_(1.0)); print(f"fill 400B of every 1200B: {ms:.3f} ms -> {2e6*400/ms/1e6:.0f} GB/s")
