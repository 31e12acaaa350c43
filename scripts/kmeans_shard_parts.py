"""Times the two per-rank phases of the sharded k-means on ONE GPU, emulating rank 0 of G:
assignment of n/G rows x all M subquantizers, ordered update of M/G subquantizers over all n rows."""
import sys

import torch

sys.path.insert(0, ".")
import reductive_b200 as rb  # noqa: E402,F401
from reductive_b200._cabi import check, lib  # noqa: E402

n, M, k, dsub = 1_000_000, 96, 256, 8
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda"); g.manual_seed(1)
cen = torch.randn((M, k, dsub), generator=g, device="cuda")
xl = torch.randn((n // G, M * dsub), generator=g, device="cuda")
Mo = M // G
xc = torch.randn((n, Mo * dsub), generator=g, device="cuda")
pitch_l, pitch_t = int(lib.rb_kmeans_code_pitch(n // G)), int(lib.rb_kmeans_code_pitch(n))
codes_l = torch.empty((M * pitch_l,), dtype=torch.uint8, device="cuda")
codes_o = torch.randint(0, 256, (Mo * pitch_t,), dtype=torch.uint8, device="cuda", generator=g)
packed = torch.empty((int(lib.rb_kmeans_packed_len(Mo, k, dsub)),), device="cuda")


def timed(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ta = timed(lambda: check(lib.rb_kmeans_assign(xl.data_ptr(), n // G, xl.stride(0), cen.data_ptr(), M, k, dsub, codes_l.data_ptr(), st)))
tu = timed(lambda: check(lib.rb_kmeans_accumulate(xc.data_ptr(), n, xc.stride(0), codes_o.data_ptr(), Mo, k, dsub, None, packed.data_ptr(), st)))
print(f"G={G}: assign {n // G} rows x {M}: {ta:.3f} ms; ordered update {Mo} subquantizers x {n} rows: {tu:.3f} ms", flush=True)
