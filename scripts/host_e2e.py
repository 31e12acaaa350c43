"""End-to-end quantize_batch through the C ABI from HOST memory, one process: pageable (plain numpy) vs pinned buffers,
one device vs all devices of the box (rb_pq_create_multi).  Rows per device: argv[1] (default 1M)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402
from reductive_b200._cabi import check, lib  # noqa: E402

rows_per_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
M, k, dsub = 30, 256, 10
d = M * dsub
q = np.random.default_rng(1).normal(size=(M, k, dsub)).astype(np.float32)
ndev = torch.cuda.device_count()
block = np.random.default_rng(2).normal(size=(100_000, d)).astype(np.float32)


def run(pq, x, c, reps=3):
    pq.quantize_batch_into(x, c)
    t0 = time.perf_counter()
    for _ in range(reps):
        pq.quantize_batch_into(x, c)
    return (time.perf_counter() - t0) / reps


for devs in ([0], list(range(ndev))) if ndev > 1 else ([0],):
    n = rows_per_dev * len(devs)
    x = np.empty((n, d), np.float32)  # pageable
    for i in range(0, n, len(block)):
        x[i:i + len(block)] = block[: min(len(block), n - i)]
    c = np.empty((n, M), np.uint8)
    pq = rb.Pq(None, q, devices=devs)
    for threads in (1, 4, 8):
        check(lib.rb_set_host_copy_threads(threads))
        t = run(pq, x, c)
        print(f"{len(devs)} device(s), pageable numpy, {threads} copy threads/device: {n / t / 1e6:.1f} Mvec/s "
              f"({n * (4 * d + M) / t / 1e9:.1f} GB/s)", flush=True)
    check(lib.rb_set_host_copy_threads(4))
    want = c.copy()
    xp = torch.empty((n, d), dtype=torch.float32, pin_memory=True)
    xp.numpy()[:] = x
    cp = torch.empty((n, M), dtype=torch.uint8, pin_memory=True)
    t = run(pq, xp.numpy(), cp.numpy())
    assert np.array_equal(cp.numpy(), want)
    print(f"{len(devs)} device(s), pinned buffers: {n / t / 1e6:.1f} Mvec/s ({n * (4 * d + M) / t / 1e9:.1f} GB/s)", flush=True)
    del x, xp, pq
print("host e2e ok")
