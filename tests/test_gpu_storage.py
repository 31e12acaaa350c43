"""Caller-side quantized storage: lookups and the fused decode + dot against the oracle.

embedding(i) = reconstruct(codes[i]) * norm[i] is bit-exact without a projection and within 1e-5 with one (the
tolerance of rb_pq_reconstruct_batch); the fused scores are f32 sums in another order than any CPU product, so they
are held to 1e-5 * sum_c |q_c| |e_c| of the float64 product (tolerance written here, as the header states it)."""
import numpy as np
import pytest

import reductive_b200 as rb
from tests.util import F, normal, orthonormal, random_codebook

pytestmark = pytest.mark.gpu

DOT_RTOL = 1e-5


def _store(n, M, k, dsub, seed, projected=False, with_norms=True):
    q = random_codebook(M, k, dsub, seed)
    proj = orthonormal(M * dsub, seed + 1) if projected else None
    rng = np.random.default_rng(seed + 2)
    codes = rng.integers(0, k, (n, M)).astype(np.uint8)
    norms = rng.uniform(0.5, 3.0, n).astype(F) if with_norms else None
    pq = rb.Pq(proj, q)
    return pq, q, proj, codes, norms


@pytest.mark.parametrize("n,M,k,dsub,projected,with_norms", [
    (5_000, 30, 256, 10, False, True), (5_000, 30, 256, 10, False, False), (3_000, 8, 16, 4, False, True),
    (4_000, 10, 256, 30, True, True), (1, 3, 4, 2, False, True)])
def test_embeddings_match_reconstruct_times_norm(oracle, n, M, k, dsub, projected, with_norms):
    pq, q, proj, codes, norms = _store(n, M, k, dsub, 11, projected, with_norms)
    store = rb.QuantizedArray(pq, codes, norms)
    assert len(store) == n and store.shape == (n, M * dsub) and store.has_norms() == with_norms
    idx = np.random.default_rng(5).integers(0, n, 777)
    idx[:3] = [0, n - 1, n // 2]
    want = oracle.qstore_embeddings(q, proj, codes, norms, idx)
    got = store.embeddings(idx)
    if projected:
        assert np.max(np.abs(got - want)) <= 1e-5 * max(1.0, np.max(np.abs(want)))
    else:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(store.embedding(int(idx[7])), got[7])
    assert store.embeddings(np.zeros((0,), np.int64)).shape == (0, M * dsub)
    with pytest.raises(IndexError):
        store.embedding(n)
    store.close()


def test_embeddings_on_device_tensors(oracle):
    import torch

    n, M, k, dsub = 20_000, 16, 256, 8
    pq, q, proj, codes, norms = _store(n, M, k, dsub, 23)
    store = rb.QuantizedArray(pq, torch.from_numpy(codes).cuda(), torch.from_numpy(norms).cuda())
    idx = torch.randint(0, n, (4_096,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    got = store.embeddings(idx).cpu().numpy()
    want = oracle.qstore_embeddings(q, proj, codes, norms, idx.cpu().numpy())
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("n,M,k,dsub,nq,projected,with_norms", [
    (10_000, 30, 256, 10, 8, False, True),    # table of 8 queries in shared memory
    (10_001, 30, 256, 10, 13, False, False),  # ragged row tile, ragged query batch
    (6_000, 96, 256, 8, 3, False, True),      # C3 geometry: one query per pass fits
    (6_000, 10, 256, 30, 5, True, True),      # rotated quantizer: the queries are rotated instead of the rows
    (3_000, 7, 2, 3, 2, False, True),         # k = 2, odd M: unaligned table
    (700, 5, 64, 4, 1, False, False)])
def test_fused_decode_dot(oracle, n, M, k, dsub, nq, projected, with_norms):
    pq, q, proj, codes, norms = _store(n, M, k, dsub, 31, projected, with_norms)
    store = rb.QuantizedArray(pq, codes, norms)
    queries = normal((nq, M * dsub), 77)
    want, mag = oracle.qstore_dot(q, proj, codes, norms, queries)
    got = store.dot(queries)
    assert got.shape == (nq, n)
    assert np.all(np.abs(got - want) <= DOT_RTOL * mag + 1e-30)


def test_fused_decode_dot_device_and_strided_queries(oracle):
    import torch

    n, M, k, dsub, nq = 50_000, 30, 256, 10, 9
    pq, q, proj, codes, norms = _store(n, M, k, dsub, 41)
    store = rb.QuantizedArray(pq, torch.from_numpy(codes).cuda(), torch.from_numpy(norms).cuda())
    big = torch.from_numpy(normal((nq, 2 * M * dsub), 3)).cuda()
    queries = big[:, ::2]  # column stride 2
    got = store.dot(queries).cpu().numpy()
    want, mag = oracle.qstore_dot(q, proj, codes, norms, queries.cpu().numpy())
    assert np.all(np.abs(got - want) <= DOT_RTOL * mag + 1e-30)
    # the argmax of the fused scores is the argmax of the exact ones wherever the exact gap exceeds the tolerance
    top = np.argmax(want, axis=1)
    gap = want[np.arange(nq), top][:, None] - want
    gap[np.arange(nq), top] = np.inf
    safe = gap.min(axis=1) > 4 * DOT_RTOL * mag.max(axis=1)
    assert np.array_equal(np.argmax(got, axis=1)[safe], top[safe])


def test_quantize_using_round_trip(oracle):
    n, M, bits, dsub = 8_000, 10, 6, 4
    x = normal((n, M * dsub), 9) * np.random.default_rng(1).uniform(0.2, 5.0, (n, 1)).astype(F)
    pq = rb.Pq(None, random_codebook(M, 1 << bits, dsub, 4) * 0.3)
    store = rb.QuantizedArray.quantize_using(pq, x, normalize=True)
    norms = np.sqrt(np.einsum("ij,ij->i", x, x)).astype(F)
    codes = oracle.quantize_batch(pq.subquantizers(), None, x / norms[:, None])
    want = oracle.qstore_embeddings(pq.subquantizers(), None, codes, norms, np.arange(n))
    assert np.array_equal(store.embeddings(np.arange(n)).view(np.uint32), want.view(np.uint32))


def test_store_rejects_bad_input():
    pq = rb.Pq(None, random_codebook(4, 16, 3, 1))
    with pytest.raises(rb.ReductivePanic):
        rb.QuantizedArray(pq, np.zeros((5, 3), np.uint8))
    with pytest.raises(Exception):  # a code that names no centroid
        rb.QuantizedArray(pq, np.full((5, 4), 16, np.uint8))
    store = rb.QuantizedArray(pq, np.zeros((5, 4), np.uint8))
    with pytest.raises(rb.ReductivePanic):
        store.dot(np.zeros((2, 11), F))
