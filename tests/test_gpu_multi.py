"""Multi-GPU paths through the C ABI from ONE process (skipped on a box with a single GPU): sharded k-means
(rb_pq_train_multi: NCCL inside the library) must give the bit-identical quantizer of a one-GPU run."""
import numpy as np
import pytest

from tests.util import normal, rows_as_initial_centroids

pytestmark = pytest.mark.gpu


def _devices():
    import torch

    return list(range(torch.cuda.device_count()))


@pytest.mark.parametrize("n,M,bits,dsub,iters", [(20_000, 8, 8, 8, 6), (5_003, 6, 6, 10, 4), (3_000, 5, 4, 4, 3),
                                                (6_001, 4, 9, 4, 3)])  # bits = 9: 4-byte codes through the exchange
def test_multi_gpu_training_is_bit_identical_to_one_gpu(oracle, n, M, bits, dsub, iters):
    import reductive_b200 as rb

    devs = _devices()
    if len(devs) < 2:
        pytest.skip("needs at least 2 GPUs")
    x = normal((n, M * dsub), 31)
    init = rows_as_initial_centroids(x, M, 1 << bits, 32)
    one, loss1 = rb.Pq.train_pq_using(M, bits, iters, 1, x, None, initial_centroids=init, return_loss=True)
    for use in (devs[:2], devs):
        many, lossn = rb.Pq.train_pq_using(M, bits, iters, 1, x, None, initial_centroids=init, return_loss=True,
                                           devices=use)
        assert np.array_equal(one.subquantizers().view(np.int32), many.subquantizers().view(np.int32)), use
        assert np.allclose(loss1, lossn, rtol=1e-4)
    want, _ = oracle.train_pq(x, M, bits, iters, 1, init, n_threads=4)
    assert np.array_equal(one.subquantizers().view(np.int32), want.view(np.int32))


def test_code_exchange_modes_agree(monkeypatch):
    """The assignments reach their owners either through the peer-mapped code matrix (stores into the owners' memory)
    or, RB_DIST_P2P=0, by ncclSend/Recv: the trained quantizer is the same bit for bit."""
    import reductive_b200 as rb

    devs = _devices()
    if len(devs) < 2:
        pytest.skip("needs at least 2 GPUs")
    n, M, bits, dsub = 30_011, 12, 8, 8
    x = normal((n, M * dsub), 51)
    init = rows_as_initial_centroids(x, M, 1 << bits, 52)
    got = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RB_DIST_P2P", mode)
        got[mode] = rb.Pq.train_pq_using(M, bits, 5, 1, x, None, initial_centroids=init, devices=devs).subquantizers()
    assert np.array_equal(got["1"].view(np.int32), got["0"].view(np.int32))
    one = rb.Pq.train_pq_using(M, bits, 5, 1, x, None, initial_centroids=init).subquantizers()
    assert np.array_equal(got["1"].view(np.int32), one.view(np.int32))


def test_multi_gpu_training_more_devices_than_subquantizers():
    import reductive_b200 as rb

    devs = _devices()
    if len(devs) < 4:
        pytest.skip("needs at least 4 GPUs")
    n, M, bits, dsub = 4_000, 2, 5, 6  # ranks 2.. own no subquantizer
    x = normal((n, M * dsub), 41)
    init = rows_as_initial_centroids(x, M, 1 << bits, 42)
    one = rb.Pq.train_pq_using(M, bits, 3, 1, x, None, initial_centroids=init)
    many = rb.Pq.train_pq_using(M, bits, 3, 1, x, None, initial_centroids=init, devices=devs[:4])
    assert np.array_equal(one.subquantizers().view(np.int32), many.subquantizers().view(np.int32))


@pytest.mark.parametrize("projected", [False, True])
def test_host_batches_split_over_devices_from_one_process(oracle, projected):
    """rb_pq_create_multi: quantize_batch / reconstruct_batch on plain (pageable) numpy arrays, rows split over every
    GPU of the box through the C ABI -- the same bits as the oracle (and hence as one GPU)."""
    import reductive_b200 as rb
    from tests.util import orthonormal, random_codebook

    devs = _devices()
    if len(devs) < 2:
        pytest.skip("needs at least 2 GPUs")
    n, M, k, dsub = 300_001, 10, 256, 10
    q = random_codebook(M, k, dsub, 3)
    r = orthonormal(M * dsub, 4) if projected else None
    x = normal((n, M * dsub), 5)
    pq = rb.Pq(r, q, devices=devs)
    codes = pq.quantize_batch(x, np.uint8)
    one = rb.Pq(r, q)
    assert np.array_equal(codes, one.quantize_batch(x, np.uint8))
    sample = np.random.default_rng(6).choice(n, 4000, replace=False)
    assert np.array_equal(codes[sample], oracle.quantize_batch(q, r, x[sample], np.uint8))
    rec = pq.reconstruct_batch(codes)
    want = one.reconstruct_batch(codes)
    if projected:
        assert np.abs(rec - want).max() <= 1e-5 * np.abs(want).max()
    else:
        assert np.array_equal(rec, want)
