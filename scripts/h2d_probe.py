"""Host<->device copy bandwidth of the box (pinned), to put the e2e number in context."""
import time
import torch
n = 2_400_000_000 // 4
h = torch.empty((n,), dtype=torch.float32, pin_memory=True)
d = torch.empty((n,), dtype=torch.float32, device="cuda")
for name, a, b in (("h2d", d, h), ("d2h", h, d)):
    a.copy_(b, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        a.copy_(b, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {n * 4 / dt / 1e9:.1f} GB/s")
