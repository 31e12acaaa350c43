"""BASELINE.json's full-size configurations on the GPU: size-independent properties plus oracle parity on a
seeded random sample of rows (the oracle finishes a few thousand rows in well under a second)."""
import numpy as np
import pytest

import reductive_b200 as rb

pytestmark = pytest.mark.gpu


def _device_data(n, d, seed):
    import torch

    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.randn((n, d), generator=g, device="cuda", dtype=torch.float32)


@pytest.mark.parametrize("name,n,M,dsub", [("C2", 2_000_000, 30, 10), ("C5-shard", 4_000_000, 16, 8),
                                            ("C3-rows", 500_000, 96, 8)])
def test_full_size_encode_decode_properties(oracle, name, n, M, dsub):
    import torch

    k, d = 256, M * dsub
    x = _device_data(n, d, 1234)
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(n, M * k, replace=False))
    picked = x[torch.as_tensor(rows, device="cuda")].cpu().numpy()
    q = np.stack([picked[m * k:(m + 1) * k, m * dsub:(m + 1) * dsub] for m in range(M)])
    pq = rb.Pq(None, q)
    for algo in (rb.ENCODE_AUTO, rb.ENCODE_EXACT):
        rb.set_encode_algo(algo)
        codes = pq.quantize_batch(x, np.uint8)
        if algo == rb.ENCODE_AUTO:
            codes_auto = codes
    rb.set_encode_algo(rb.ENCODE_AUTO)
    # (1) both encode kernels agree on every one of the n*M codes
    assert torch.equal(codes, codes_auto), f"{name}: {(codes != codes_auto).sum().item()} codes differ between kernels"
    # (2) oracle parity on a seeded sample of rows
    sample = np.sort(rng.choice(n, 4096, replace=False))
    xs = x[torch.as_tensor(sample, device="cuda")].cpu().numpy()
    want = oracle.quantize_batch(q, None, xs, np.uint8, n_threads=8)
    got = codes[torch.as_tensor(sample, device="cuda")].cpu().numpy()
    assert np.array_equal(got, want), f"{name}: {(got != want).sum()} sampled codes differ from the oracle"
    # (3) reconstruct == an independent gather (torch advanced indexing) on all rows, bit for bit
    rec = pq.reconstruct_batch(codes)
    qt = torch.from_numpy(q).cuda()
    for m in range(0, M, max(1, M // 6)):
        ref = qt[m][codes[:, m].long()]
        assert torch.equal(rec[:, m * dsub:(m + 1) * dsub], ref), f"{name}: gather mismatch in subquantizer {m}"
    # (4) idempotence: a reconstruction encodes back to the codes it came from (distinct centroids)
    again = pq.quantize_batch(rec, np.uint8)
    assert torch.equal(again, codes), f"{name}: {(again != codes).sum().item()} codes change on re-encoding"
    # (5) each row's own code is at least as close as the runner-up: distance to the decoded vector is minimal
    #     for the sample (direct FP64 check, independent of both kernels' arithmetic)
    d2 = ((xs[:, None, :dsub].astype(np.float64) - q[0][None].astype(np.float64)) ** 2).sum(-1)
    slack = d2[np.arange(len(xs)), got[:, 0]] - d2.min(1)
    assert slack.max() <= 1e-4
