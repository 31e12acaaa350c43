"""Pure-Python, exactly-rounded model of the same arithmetic as oracle.c — for SMALL cases only.

TEST INFRASTRUCTURE ONLY.  It exists so that the C oracle (and its AVX2 micro-kernel) is itself
checked by an independent implementation: every float32 operation here is a single correctly
rounded IEEE operation (fma through exact rational arithmetic), written as plain loops.

Follows: src/linalg.rs:150-180 (distances), src/kmeans.rs:133-198,330-360 (argmin, update, mse).
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np

F = np.float32


def _round_f32(fr: Fraction) -> np.float32:
    """Round an exact rational to the nearest float32, ties to even."""
    if fr == 0:
        return F(0.0)
    f = F(float(fr))
    if not np.isfinite(f):
        return f
    cands = [np.nextafter(f, F(-np.inf)), f, np.nextafter(f, F(np.inf))]
    best, best_err = None, None
    for c in cands:
        if not np.isfinite(c):
            continue
        err = abs(Fraction(float(c)) - fr)
        if best is None or err < best_err:
            best, best_err = c, err
        elif err == best_err:
            # tie: even mantissa wins
            if (int(np.array(c).view(np.uint32)) & 1) == 0:
                best = c
    return F(best)


def fma(a, b, c) -> np.float32:
    a, b, c = F(a), F(b), F(c)
    if not (np.isfinite(a) and np.isfinite(b) and np.isfinite(c)):
        return F(a * b + c)
    return _round_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def unrolled_dot(x, y) -> np.float32:
    x = np.asarray(x, F)
    y = np.asarray(y, F)
    n = len(x)
    p = [F(0)] * 8
    i = 0
    while i + 8 <= n:
        for l in range(8):
            p[l] = F(p[l] + F(x[i + l] * y[i + l]))
        i += 8
    s = F(0)
    s = F(s + F(p[0] + p[4]))
    s = F(s + F(p[1] + p[5]))
    s = F(s + F(p[2] + p[6]))
    s = F(s + F(p[3] + p[7]))
    while i < n:
        s = F(s + F(x[i] * y[i]))
        i += 1
    return s


def gemm_elem(arow, bcol) -> np.float32:
    """One C element of the matrixmultiply model: FMA chain inside kc=256 blocks, blocks added."""
    k = len(arow)
    total = None
    for kb in range(0, max(k, 1), 256):
        acc = F(0)
        for t in range(kb, min(kb + 256, k)):
            acc = fma(arow[t], bcol[t], acc)
        total = acc if total is None else F(total + acc)
    return F(0) if total is None else total


def sqdist_batch(x, c):
    x = np.asarray(x, F)
    c = np.asarray(c, F)
    n, k = x.shape[0], c.shape[0]
    xs = [unrolled_dot(x[i], x[i]) for i in range(n)]
    cs = [unrolled_dot(c[j], c[j]) for j in range(k)]
    out = np.zeros((n, k), F)
    for i in range(n):
        for j in range(k):
            dp = gemm_elem(x[i], c[j])
            out[i, j] = F(F(xs[i] + cs[j]) - F(dp + dp))
    return out


def of_less(a, b) -> bool:
    a, b = float(a), float(b)
    if a < b:
        return True
    return (b != b) and (a == a)


def argmin_first(row) -> int:
    best, bv = 0, row[0]
    for j in range(1, len(row)):
        if of_less(row[j], bv):
            best, bv = j, row[j]
    return best


def cluster_assignments(c, x):
    d = sqdist_batch(x, c)
    return np.array([argmin_first(d[i]) for i in range(d.shape[0])], np.uint64)


def update_centroids(c, x, assign):
    c = np.zeros_like(np.asarray(c, F))
    x = np.asarray(x, F)
    counts = np.zeros((c.shape[0],), F)
    for i in range(x.shape[0]):
        a = int(assign[i])
        for t in range(c.shape[1]):
            c[a, t] = F(c[a, t] + x[i, t])
        counts[a] = F(counts[a] + F(1))
    for j in range(c.shape[0]):
        if counts[j] > 0:
            for t in range(c.shape[1]):
                c[j, t] = F(c[j, t] / counts[j])
    return c


def mean_squared_error(c, x, assign):
    c = np.asarray(c, F)
    x = np.asarray(x, F)
    sse = F(0)
    for i in range(x.shape[0]):
        for t in range(x.shape[1]):
            v = F(c[int(assign[i]), t] - x[i, t])
            sse = F(sse + F(v * v))
    return F(sse / F(x.shape[0] * x.shape[1]))
