import sys, time, torch
sys.path.insert(0, ".")
import reductive_b200 as rb
from reductive_b200.dist import cuda_accumulate, cuda_assign, cuda_finalize
M, k, dsub = 96, 256, 8
n = 32768
g = torch.Generator(device="cuda"); g.manual_seed(77)
x = torch.randn((n, M * dsub), generator=g, device="cuda")
cen = torch.randn((M, k, dsub), generator=g, device="cuda")
packed = torch.empty((M * k * dsub + M * k + M,), device="cuda")
for _ in range(3):
    codes = cuda_assign(x, cen)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    codes = cuda_assign(x, cen)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"assign: host {1e6*(t1-t0)/20:.0f} us per call, with drain {1e6*(t2-t0)/20:.0f} us")
t0 = time.perf_counter()
for _ in range(20):
    cuda_accumulate(x, cen, codes, None, packed)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"accumulate: host {1e6*(t1-t0)/20:.0f} us per call, with drain {1e6*(t2-t0)/20:.0f} us")
pq = rb.Pq(None, cen.cpu().numpy())
out = torch.empty((n, M), dtype=torch.uint8, device="cuda")
pq.quantize_batch_into(x, out); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    pq.quantize_batch_into(x, out)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"quantize_batch: host {1e6*(t1-t0)/20:.0f} us per call, with drain {1e6*(t2-t0)/20:.0f} us")
