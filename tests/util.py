"""Shared generators for the parity tests (seeded, so the oracle and the CUDA path see identical inputs)."""
import numpy as np

F = np.float32


def normal(shape, seed):
    return np.random.default_rng(seed).normal(size=shape).astype(F)


def random_codebook(M, k, dsub, seed):
    return normal((M, k, dsub), seed)


def rows_as_initial_centroids(x, M, k, seed, n_attempts=1):
    """[n_attempts, M, k, dsub]: k distinct instance rows per subquantizer and attempt."""
    rng = np.random.default_rng(seed)
    n, d = x.shape
    dsub = d // M
    out = np.empty((n_attempts, M, k, dsub), F)
    for a in range(n_attempts):
        for m in range(M):
            out[a, m] = x[rng.choice(n, k, replace=False)][:, m * dsub:(m + 1) * dsub]
    return out


def orthonormal(d, seed):
    q, _ = np.linalg.qr(np.random.default_rng(seed).normal(size=(d, d)))
    return np.ascontiguousarray(q, F)


def near_tie_rows(q, n, seed):
    """Rows whose sub-vectors sit on (or a few ulp off) the bisector of two centroids, plus exact centroid
    copies: the adversarial set for tie-breaking / rescoring."""
    rng = np.random.default_rng(seed)
    M, k, dsub = q.shape
    x = np.empty((n, M * dsub), F)
    for m in range(M):
        a = rng.integers(0, k, n)
        b = (a + 1 + rng.integers(0, k - 1, n)) % k
        mid = (q[m, a].astype(np.float64) + q[m, b].astype(np.float64)) / 2
        sub = mid.astype(F)
        kind = rng.integers(0, 4, n)
        sub = np.where((kind == 1)[:, None], np.nextafter(sub, q[m, a]), sub)
        sub = np.where((kind == 2)[:, None], np.nextafter(sub, q[m, b]), sub)
        sub = np.where((kind == 3)[:, None], q[m, a], sub)
        x[:, m * dsub:(m + 1) * dsub] = sub
    return x
