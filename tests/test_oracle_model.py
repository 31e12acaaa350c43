"""The C oracle against an independent exactly-rounded pure-Python model (oracle/pymodel.py),
its scalar build against its AVX2 build, and the semantic edge cases the reference's tests leave
unpinned (ties, NaN, empty clusters, truncating cast, strides, the kc=256 split)."""
import numpy as np
import pytest

from oracle import pymodel

F = np.float32


def _bits(a):
    return np.ascontiguousarray(a, F).view(np.uint32)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 10, 15, 16, 17, 30, 33])
def test_unrolled_dot_matches_model(oracle, n):
    rng = np.random.default_rng(n)
    x, y = rng.normal(size=n).astype(F), rng.normal(size=n).astype(F)
    assert _bits(oracle.unrolled_dot(x, y)) == _bits(pymodel.unrolled_dot(x, y))


@pytest.mark.parametrize("dsub,k,n", [(10, 16, 9), (8, 24, 5), (30, 17, 6), (3, 5, 7), (1, 4, 4)])
def test_sqdist_matches_model_bit_exact(oracle, dsub, k, n):
    rng = np.random.default_rng(dsub * 100 + k)
    x = rng.normal(size=(n, dsub)).astype(F)
    c = rng.normal(size=(k, dsub)).astype(F)
    assert np.array_equal(_bits(oracle.sqdist_batch(x, c)), _bits(pymodel.sqdist_batch(x, c)))


def test_sgemm_kc_split_matches_model(oracle):
    # d = 300 > kc = 256: C = chain(0..255) + chain(256..299)   
    rng = np.random.default_rng(3)
    a = rng.normal(size=(3, 300)).astype(F)
    b = rng.normal(size=(300, 5)).astype(F)
    got = oracle.sgemm(a, b)
    want = np.array([[pymodel.gemm_elem(a[i], b[:, j]) for j in range(5)] for i in range(3)], F)
    assert np.array_equal(_bits(got), _bits(want))
    # and a transposed-B view (pq.rs:324 passes projection.t())
    bt = np.ascontiguousarray(b.T)
    assert np.array_equal(_bits(oracle.sgemm(a, bt.T)), _bits(want))


def test_update_and_mse_match_model(oracle):
    rng = np.random.default_rng(5)
    x = rng.normal(size=(40, 6)).astype(F)
    c = rng.normal(size=(7, 6)).astype(F)
    a = oracle.cluster_assignments(c, x)
    assert np.array_equal(a, pymodel.cluster_assignments(c, x))
    c2 = oracle.update_centroids(c, x, a)
    assert np.array_equal(_bits(c2), _bits(pymodel.update_centroids(c, x, a)))
    assert _bits(oracle.mean_squared_error(c2, x, a)) == _bits(pymodel.mean_squared_error(c2, x, a))


def test_scalar_and_avx2_builds_agree(oracle, oracle_scalar):
    rng = np.random.default_rng(11)
    M, k, dsub, n = 6, 64, 10, 3000
    q = rng.normal(size=(M, k, dsub)).astype(F)
    x = rng.normal(size=(n, M * dsub)).astype(F)
    a = oracle.quantize_batch(q, None, x, np.uint8, n_threads=3)
    b = oracle_scalar.quantize_batch(q, None, x, np.uint8)
    assert np.array_equal(a, b)
    d1 = oracle.sqdist_batch(x[:257, :dsub], q[0])
    d2 = oracle_scalar.sqdist_batch(x[:257, :dsub], q[0])
    assert np.array_equal(_bits(d1), _bits(d2))
    r = np.linalg.qr(rng.normal(size=(60, 60)))[0].astype(F)
    assert np.array_equal(oracle.quantize_batch(q, r, x, np.uint8, n_threads=2),
                          oracle_scalar.quantize_batch(q, r, x, np.uint8))


def test_first_index_wins_ties(oracle):
    # duplicated centroids: exact ties -> lowest index (min_by_key keeps the first minimum)
    c = np.array([[1., 1.], [0., 0.], [1., 1.], [0., 0.]], F)
    x = np.array([[0.1, 0.], [0.9, 1.], [0.5, 0.5]], F)
    assert oracle.cluster_assignments(c, x).tolist() == [1, 0, 0]
    assert [oracle.cluster_assignment(c, v) for v in x] == [1, 0, 0]


def test_nan_ranks_largest(oracle):
    c = np.array([[np.nan, 0.], [1., 1.], [np.nan, np.nan]], F)
    x = np.array([[0., 0.], [5., 5.]], F)
    assert oracle.cluster_assignments(c, x).tolist() == [1, 1]
    # all NaN: every comparison is Equal -> index 0
    call = np.full((3, 2), np.nan, F)
    assert oracle.cluster_assignments(call, x).tolist() == [0, 0]
    # NaN instance: all distances NaN -> 0
    assert oracle.cluster_assignments(c[1:2].repeat(3, 0), np.array([[np.nan, 1.]], F)).tolist() == [0]


def test_empty_cluster_stays_zero(oracle):  # kmeans.rs:181,194
    c = np.ones((3, 2), F)
    x = np.array([[1., 2.], [3., 4.]], F)
    got = oracle.update_centroids(c, x, np.array([0, 0], np.uint64))
    assert np.array_equal(got, np.array([[2., 3.], [0., 0.], [0., 0.]], F))


def test_batch_cast_truncates(oracle):  # primitives.rs:100: `as` cast, no overflow check in the batch path
    rng = np.random.default_rng(2)
    q = rng.normal(size=(1, 300, 4)).astype(F)
    x = q[0, 256:300].copy()  # nearest centroid of row i is 256 + i
    wide = oracle.quantize_batch(q, None, x, np.uint32)
    narrow = oracle.quantize_batch(q, None, x, np.uint8)
    assert np.array_equal(wide[:, 0], np.arange(256, 300))
    assert np.array_equal(narrow[:, 0], (np.arange(256, 300) & 0xFF).astype(np.uint8))


def test_strided_input_same_codes_on_exact_data(oracle):
    # column-major input takes the sequential-dot norm path; on exactly representable data
    # every summation order agrees
    rng = np.random.default_rng(4)
    q = rng.integers(-3, 4, size=(3, 8, 4)).astype(F)
    x = rng.integers(-3, 4, size=(50, 12)).astype(F)
    assert np.array_equal(oracle.quantize_batch(q, None, x, np.uint8),
                          oracle.quantize_batch(q, None, np.asfortranarray(x), np.uint8))


def test_vector_and_batch_agree_on_exact_data(oracle):
    rng = np.random.default_rng(6)
    q = rng.integers(-4, 5, size=(4, 16, 5)).astype(F)
    x = rng.integers(-4, 5, size=(30, 20)).astype(F)
    r = np.eye(20, dtype=F)[rng.permutation(20)]
    for proj in (None, r):
        b = oracle.quantize_batch(q, proj, x, np.uint16)
        for i in range(30):
            assert np.array_equal(oracle.quantize_vector(q, proj, x[i], np.uint16), b[i])
        rec = oracle.reconstruct_batch(q, proj, b)
        for i in range(30):
            assert np.array_equal(oracle.reconstruct(q, proj, b[i]), rec[i])


def test_reconstruct_rejects_bad_code(oracle):
    q = np.zeros((2, 4, 3), F)
    with pytest.raises(IndexError):
        oracle.reconstruct_batch(q, None, np.array([[0, 4]], np.uint8))


def test_multi_attempt_picks_lowest_loss(oracle):
    rng = np.random.default_rng(8)
    x = rng.normal(size=(300, 8)).astype(F)
    M, bits = 2, 3
    init = np.stack([np.stack([x[rng.choice(300, 8, replace=False)][:, m * 4:(m + 1) * 4] for m in range(M)])
                     for _ in range(3)])
    q, loss = oracle.train_pq(x, M, bits, 4, 3, init)
    singles = [oracle.train_pq(x, M, bits, 4, 1, init[a:a + 1]) for a in range(3)]
    for m in range(M):
        best = min(range(3), key=lambda a: (singles[a][1][m], a))
        assert np.array_equal(q[m], singles[best][0][m]) and loss[m] == singles[best][1][m]
