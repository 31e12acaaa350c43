"""The CPU oracle against EVERY known-answer vector the reference's own tests hold for the hot path.

Each test names the reference test it restates (file:line under /root/reference).
"""
import numpy as np
import pytest

F = np.float32


# ---- src/pq/pq.rs:378-407 fixtures -------------------------------------------------------------
def _test_vectors():
    return np.array([[0., 2., 0., -0.5, 0., 0.],
                     [1., -0.2, 0., 0.5, 0.5, 0.],
                     [-0.2, 0.2, 0., 0., -2., 0.],
                     [1., 0.2, 0., 0., -2., 0.]], F)


def _test_quantizations():
    return np.array([[1, 1], [0, 1], [1, 0], [0, 0]], np.uint64)


def _test_reconstructions():
    return np.array([[0., 1., 0., 0., 1., 0.],
                     [1., 0., 0., 0., 1., 0.],
                     [0., 1., 0., 1., -1., 0.],
                     [1., 0., 0., 1., -1., 0.]], F)


def _test_quantizers():
    return np.array([[[1., 0., 0.], [0., 1., 0.]], [[1., -1., 0.], [0., 1., 0.]]], F)


def test_quantize_batch_with_predefined_codebook(oracle):  # pq.rs:409-417
    got = oracle.quantize_batch(_test_quantizers(), None, _test_vectors(), np.uint64)
    assert np.array_equal(got, _test_quantizations())


def test_quantize_with_predefined_codebook(oracle):  # pq.rs:419-429
    for v, q in zip(_test_vectors(), _test_quantizations()):
        assert np.array_equal(oracle.quantize_vector(_test_quantizers(), None, v, np.uint64), q)


def test_reconstruct_batch_with_predefined_codebook(oracle):  # pq.rs:471-478
    got = oracle.reconstruct_batch(_test_quantizers(), None, _test_quantizations())
    assert np.array_equal(got, _test_reconstructions())


def test_reconstruct_with_predefined_codebook(oracle):  # pq.rs:480-490
    for q, r in zip(_test_quantizations(), _test_reconstructions()):
        assert np.array_equal(oracle.reconstruct(_test_quantizers(), None, q), r)


def test_quantize_with_type(oracle):  # pq.rs:442-450: u8 with 256 centroids is fine
    rng = np.random.default_rng(0)
    q = rng.random((1, 256, 10), F)
    oracle.quantize_vector(q, None, rng.random((10,), F), np.uint8)


def test_quantize_with_too_narrow_type(oracle):  # pq.rs:452-461: u8 with 257 centroids panics
    rng = np.random.default_rng(0)
    q = rng.random((1, 257, 10), F)
    with pytest.raises(OverflowError):
        oracle.quantize_vector(q, None, rng.random((10,), F), np.uint8)


# ---- src/kmeans.rs:380-435, 504-519 ------------------------------------------------------------
def test_correct_cluster_assignments(oracle):  # kmeans.rs:380-400
    centroids = np.array([[0.5, 0., 0.], [0., -1., 0.], [0., 0., 1.], [0., 1., 1.]], F)
    instances = np.array([[0., 0.5, 0.], [0., 0., 2.], [1., 0., 0.], [0., 0., 1.],
                          [0., -2., 0.], [0., 0.7, 0.7], [0., 0., 0.]], F)
    want = np.array([0, 2, 0, 2, 1, 3, 0], np.uint64)
    assert np.array_equal(oracle.cluster_assignments(centroids, instances), want)
    # Axis(1): instances.t() of the transposed array is the same logical matrix (kmeans.rs:398)
    assert np.array_equal(oracle.cluster_assignments(centroids, np.asfortranarray(instances)), want)
    for x, a in zip(instances, want):
        assert oracle.cluster_assignment(centroids, x) == a


def test_correct_update_centroids(oracle):  # kmeans.rs:402-435
    centroids = np.array([[1., 0., 0.], [0., 1., 0.], [0., 0., 1.]], F)
    instances = np.array([[-1., -1., 0.], [1., 1., 0.], [-2., -1., 0.],
                          [0., 0., 0.], [0., 0., 1.], [0., 0., 2.]], F)
    assignments = np.array([1, 0, 1, 0, 2, 2], np.uint64)
    got = oracle.update_centroids(centroids, instances, assignments)
    assert np.array_equal(got, np.array([[0.5, 0.5, 0.], [-1.5, -1., 0.], [0., 0., 1.5]], F))


def test_correct_mean_squared_error(oracle):  # kmeans.rs:504-519
    centroids = np.array([[-1., 2., 0.], [0., -1., 1.]], F)
    instances = np.array([[-1., 1., 1.], [0., 1., 0.]], F)
    mse = oracle.mean_squared_error(centroids, instances, np.array([1, 0], np.uint64))
    assert mse == F(7.) / F(6.)


def test_k_means_3(oracle):  # kmeans.rs:459-479 (statistical: three tight Gaussian blobs are recovered)
    rng = np.random.default_rng(1234)
    centers = np.array([[0., 0.], [1., 0.], [1., 1.]], F)
    x = np.concatenate([c + rng.normal(0, 0.01, (11, 2)).astype(F) for c in centers]).astype(F)
    init = x[[0, 11, 22]]  # one instance per blob, as a fixed seed gives in the reference
    c, _ = oracle.kmeans_with_centroids(x, init, 10)
    got = sorted(map(tuple, np.rint(c).astype(int).tolist()))
    assert got == [(0, 0), (1, 0), (1, 1)]


# ---- src/linalg.rs:291-313 ---------------------------------------------------------------------
def test_squared_euclidean_distance_ix1_ix1(oracle):  # linalg.rs:291-296
    a = np.array([1., 2., 3.], F)
    b = np.array([0., 2., 0.], F)
    got = oracle.unrolled_dot(a, a) + oracle.unrolled_dot(b, b) - (oracle.unrolled_dot(a, b) * F(2))
    assert got == F(10)


def test_squared_euclidean_distances_ix1_ix2(oracle):  # linalg.rs:298-303
    a = np.array([1., 2., 3.], F)
    b = np.array([[2., 0., 0.], [0., 2., 0.], [0., 0., 2.]], F)
    assert np.array_equal(oracle.sqdist_vec(a, b), np.array([14., 10., 6.], F))


def test_squared_euclidean_distances_ix2_ix2(oracle):  # linalg.rs:305-313
    a = np.array([[1., 2., 3.], [3., 2., 1.]], F)
    b = np.array([[2., 0., 0.], [0., 2., 0.], [0., 0., 2.]], F)
    assert np.array_equal(oracle.sqdist_batch(a, b), np.array([[14., 10., 6.], [6., 10., 14.]], F))


# ---- src/pq/pq.rs:63-100 -----------------------------------------------------------------------
def test_check_quantizer_invariants(oracle):
    ok = oracle.check_quantizer_invariants
    assert ok(10, 7, 10, 1, 256, 20) == (0, 0)
    assert ok(0, 7, 10, 1, 256, 20) == (5, 20)      # NSubquantizersOutsideRange
    assert ok(21, 7, 10, 1, 256, 20) == (5, 20)
    assert ok(10, 0, 10, 1, 256, 20) == (3, 8)      # IncorrectNSubquantizerBits, max = floor(log2 256)
    assert ok(10, 9, 10, 1, 256, 20) == (3, 8)
    assert ok(10, 8, 10, 1, 256, 20) == (0, 0)
    assert ok(10, 8, 10, 1, 255, 20)[0] == 3
    assert ok(3, 7, 10, 1, 256, 20)[0] == 4         # IncorrectNumberSubquantizers
    assert ok(10, 7, 0, 1, 256, 20)[0] == 2         # IncorrectNIterations
    assert ok(10, 7, 10, 0, 256, 20)[0] == 1        # IncorrectNAttempts
    assert ok(10, 1, 10, 1, 0, 20)[0] == 3          # zero rows: log2(0) -> 0 bits allowed


# ---- statistical end-to-end, src/pq/pq.rs:431-440 ----------------------------------------------
def test_quantize_with_pq_statistical(oracle):
    rng = np.random.default_rng(42)
    x = rng.random((256, 20), F)
    M, bits = 10, 7
    k = 1 << bits
    init = np.stack([x[rng.choice(256, k, replace=False)][:, m * 2:(m + 1) * 2] for m in range(M)])[None]
    q, _ = oracle.train_pq(x, M, bits, 10, 1, init, n_threads=2)
    codes = oracle.quantize_batch(q, None, x, np.uint8)
    rec = oracle.reconstruct_batch(q, None, codes)
    loss = np.mean(np.sqrt(((x - rec) ** 2).sum(1)))
    assert loss < 0.08


# ---- OPQ-training restatement (oracle.py: covariance / create_projection_matrix / opq_train_iteration) ----------
def test_covariance_known_answer(oracle):  # linalg.rs:253-261
    x = np.array([[0., 2.], [1., 1.], [2., 0.]], F)
    assert np.array_equal(oracle.covariance(x), np.array([[1., -1.], [-1., 1.]], F))
    # the transposed view with the other axis (linalg.rs:259-260) is the same computation on x
    assert np.array_equal(oracle.covariance(np.ascontiguousarray(x.T).T), np.array([[1., -1.], [-1., 1.]], F))


def test_covariance_agrees_with_float64(oracle):
    rng = np.random.default_rng(3)
    x = (rng.normal(size=(4_000, 24)) * np.linspace(0.5, 3, 24) + np.linspace(-1, 1, 24)).astype(F)
    c = x.astype(np.float64) - x.mean(0, dtype=np.float64)
    want = c.T @ c / (len(x) - 1)
    got = oracle.covariance(x)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_projection_matrix_is_an_eigenbasis_permutation(oracle):  # opq.rs:103-136
    """create_projection_matrix: the eigenvectors of the covariance, columns permuted by the eigenvalue bucketing
    (opq.rs:212-273): orthonormal, and it diagonalises the covariance."""
    rng = np.random.default_rng(5)
    x = (rng.normal(size=(3_000, 12)) @ rng.normal(size=(12, 12))).astype(F)
    p = oracle.create_projection_matrix(x, 4)
    assert p.shape == (12, 12) and p.dtype == F
    assert np.abs(p.T @ p - np.eye(12)).max() < 1e-5
    rot = p.astype(np.float64).T @ oracle.covariance(x).astype(np.float64) @ p.astype(np.float64)
    off = rot - np.diag(np.diag(rot))
    assert np.abs(off).max() < 1e-4 * np.abs(np.diag(rot)).max()
    # buckets of 3 consecutive columns: their eigenvalue products are balanced (what bucket_eigenvalues optimises)
    logs = np.log(np.diag(rot)).reshape(4, 3).sum(1)
    assert logs.max() - logs.min() < np.log(np.diag(rot)).max() - np.log(np.diag(rot)).min()


def test_opq_train_iteration_is_a_rotation_and_does_not_increase_the_error(oracle):  # opq.rs:161-189
    rng = np.random.default_rng(7)
    M, k, dsub, n = 3, 8, 2, 1_500
    x = (rng.normal(size=(n, M * dsub)) @ rng.normal(size=(M * dsub, M * dsub))).astype(F)
    r = oracle.create_projection_matrix(x, M)
    rx = oracle.sgemm(x, r)
    c = np.stack([rx[rng.choice(n, k, replace=False)][:, m * dsub:(m + 1) * dsub] for m in range(M)])

    def err(r_, c_):
        rx_ = oracle.sgemm(x, r_)
        rec = oracle.reconstruct_batch(c_, None, oracle.quantize_batch(c_, None, rx_, np.uint32))
        return float(((rx_ - rec) ** 2).sum())

    e0 = err(r, c)
    for _ in range(3):
        r, c, xty = oracle.opq_train_iteration(r, c, x)
        assert np.abs(r.T @ r - np.eye(M * dsub)).max() < 1e-4 and xty.shape == (M * dsub, M * dsub)
    assert err(r, c) <= e0 * 1.0001


def test_qstore_restatement_is_reconstruct_times_norm(oracle):
    """The storage layer's oracle (parity unpinned: outside the reference tree) is reconstruct_batch x norm, and its
    exact scores are the float64 products of those embeddings."""
    rng = np.random.default_rng(9)
    q = rng.normal(size=(3, 16, 4)).astype(F)
    codes = rng.integers(0, 16, (50, 3)).astype(np.uint8)
    norms = rng.uniform(0.5, 2, 50).astype(F)
    idx = np.array([0, 49, 7, 7])
    e = oracle.qstore_embeddings(q, None, codes, norms, idx)
    assert np.array_equal(e, oracle.reconstruct_batch(q, None, codes[idx]) * norms[idx][:, None])
    queries = rng.normal(size=(2, 12)).astype(F)
    s, mag = oracle.qstore_dot(q, None, codes, norms, queries)
    full = oracle.qstore_embeddings(q, None, codes, norms, np.arange(50)).astype(np.float64)
    assert np.allclose(s, queries.astype(np.float64) @ full.T) and np.all(mag >= np.abs(s) - 1e-12)
