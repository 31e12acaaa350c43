"""Multi-GPU paths through the C ABI from ONE process (skipped on a box with a single GPU): sharded k-means
(rb_pq_train_multi: NCCL inside the library) must give the bit-identical quantizer of a one-GPU run."""
import numpy as np
import pytest

from tests.util import normal, rows_as_initial_centroids

pytestmark = pytest.mark.gpu


def _devices():
    import torch

    return list(range(torch.cuda.device_count()))


@pytest.mark.parametrize("n,M,bits,dsub,iters", [(20_000, 8, 8, 8, 6), (5_003, 6, 6, 10, 4), (3_000, 5, 4, 4, 3)])
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
