"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on identical seeded
inputs.  Bit-exact for codes and Pq reconstructions (integer / copy work), bit-exact as well for the projected
paths (the FP32 rotation keeps the reference's accumulation order), <= 1e-4 relative for trained centroids.

Every test runs once per encode algorithm (exact SIMT kernel, and AUTO = tensor path where the shape allows)."""
import numpy as np
import pytest

import reductive_b200 as rb
from tests.util import F, near_tie_rows, normal, orthonormal, random_codebook, rows_as_initial_centroids

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["exact", "auto"])
def algo(request):
    rb.set_encode_algo(rb.ENCODE_EXACT if request.param == "exact" else rb.ENCODE_AUTO)
    yield request.param
    rb.set_encode_algo(rb.ENCODE_AUTO)


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    return torch


# ---- the reference's own known-answer vectors, through the CUDA path ----------------------------------
def _test_vectors():
    return np.array([[0., 2., 0., -0.5, 0., 0.], [1., -0.2, 0., 0.5, 0.5, 0.],
                     [-0.2, 0.2, 0., 0., -2., 0.], [1., 0.2, 0., 0., -2., 0.]], F)


_QUANT = np.array([[1, 1], [0, 1], [1, 0], [0, 0]])
_RECON = np.array([[0., 1., 0., 0., 1., 0.], [1., 0., 0., 0., 1., 0.],
                   [0., 1., 0., 1., -1., 0.], [1., 0., 0., 1., -1., 0.]], F)
_QUANTIZERS = np.array([[[1., 0., 0.], [0., 1., 0.]], [[1., -1., 0.], [0., 1., 0.]]], F)


def test_quantize_batch_with_predefined_codebook(algo):  # pq.rs:409-417
    pq = rb.Pq(None, _QUANTIZERS)
    assert np.array_equal(pq.quantize_batch(_test_vectors(), np.uint64), _QUANT)


def test_quantize_with_predefined_codebook():  # pq.rs:419-429
    pq = rb.Pq(None, _QUANTIZERS)
    for v, q in zip(_test_vectors(), _QUANT):
        assert np.array_equal(pq.quantize_vector(v, np.uint64), q)


def test_reconstruct_batch_with_predefined_codebook():  # pq.rs:471-478
    pq = rb.Pq(None, _QUANTIZERS)
    assert np.array_equal(pq.reconstruct_batch(_QUANT.astype(np.uint64)), _RECON)


def test_reconstruct_with_predefined_codebook():  # pq.rs:480-490
    pq = rb.Pq(None, _QUANTIZERS)
    for q, r in zip(_QUANT.astype(np.uint64), _RECON):
        assert np.array_equal(pq.reconstruct(q), r)


def test_quantizer_lens():  # pq.rs:463-469
    pq = rb.Pq(None, _QUANTIZERS)
    assert pq.quantized_len() == 2 and pq.reconstructed_len() == 6 and pq.n_quantizer_centroids() == 2
    assert np.array_equal(pq.subquantizers(), _QUANTIZERS) and pq.projection() is None


def test_quantize_with_type_and_too_narrow_type():  # pq.rs:442-461
    q = np.random.default_rng(0).random((1, 256, 10), F)
    rb.Pq(None, q).quantize_vector(np.random.default_rng(1).random((10,), F), np.uint8)
    q257 = np.random.default_rng(0).random((1, 257, 10), F)
    with pytest.raises(rb.ReductivePanic):
        rb.Pq(None, q257).quantize_vector(np.random.default_rng(1).random((10,), F), np.uint8)


def test_correct_cluster_assignments_and_update(torch_cuda):  # kmeans.rs:380-435
    torch = torch_cuda
    centroids = np.array([[0.5, 0., 0.], [0., -1., 0.], [0., 0., 1.], [0., 1., 1.]], F)
    instances = np.array([[0., 0.5, 0.], [0., 0., 2.], [1., 0., 0.], [0., 0., 1.],
                          [0., -2., 0.], [0., 0.7, 0.7], [0., 0., 0.]], F)
    codes = rb.Pq(None, centroids[None]).quantize_batch(instances, np.uint64)
    assert codes[:, 0].tolist() == [0, 2, 0, 2, 1, 3, 0]
    # update_centroids known answer (kmeans.rs:402-435) through one kmeans_iteration from centroids that
    # reproduce the fixture's assignments [1, 0, 1, 0, 2, 2]
    inst = np.array([[-1., -1., 0.], [1., 1., 0.], [-2., -1., 0.], [0., 0., 0.], [0., 0., 1.], [0., 0., 2.]], F)
    cen = torch.tensor([[0.6, 0.6, 0.], [-1.4, -1., 0.], [0., 0., 1.4]], device="cuda")
    loss = rb.kmeans_iteration(torch.from_numpy(inst).cuda(), cen)
    assert np.array_equal(cen.cpu().numpy(), np.array([[0.5, 0.5, 0.], [-1.5, -1., 0.], [0., 0., 1.5]], F))
    assert abs(loss - (0.5 + 0.5 + 0.25 + 0.25 + 0.25 + 0.25) / 18) < 1e-6


# ---- random parity against the oracle -------------------------------------------------------------------
SHAPES = [  # (n, M, k, dsub)
    (10_000, 10, 256, 30),   # BASELINE config C1
    (20_000, 30, 256, 10),   # C2 geometry
    (8_192, 96, 256, 8),     # C3 geometry
    (16_384, 16, 256, 8),    # C5 geometry
    (100, 16, 16, 8),        # benches/pq.rs shape
    (6_000, 6, 128, 10),     # 7-bit codebook: tensor path with padded columns
    (5_000, 4, 200, 8),      # k not a power of two, padded
    (3_001, 4, 64, 7),       # odd dsub, ragged row count
    (1_234, 3, 100, 11),     # dsub outside the templated set -> generic kernel; k not a multiple of 16
    (777, 5, 300, 4),        # k > 256
    (513, 1, 32, 300),       # dsub > 256: the kc = 256 block split
    (1, 2, 8, 5),            # single row
]


@pytest.mark.parametrize("n,M,k,dsub", SHAPES)
def test_quantize_and_reconstruct_batch_bit_exact(oracle, algo, n, M, k, dsub):
    x = normal((n, M * dsub), 100 + n)
    q = random_codebook(M, k, dsub, 200 + n)
    dt = np.uint8 if k <= 256 else np.uint16
    pq = rb.Pq(None, q)
    codes = pq.quantize_batch(x, dt)
    want = oracle.quantize_batch(q, None, x, dt, n_threads=8)
    assert np.array_equal(codes, want), f"{(codes != want).sum()} of {codes.size} codes differ"
    rec = pq.reconstruct_batch(codes)
    assert np.array_equal(rec.view(np.uint32), oracle.reconstruct_batch(q, None, want).view(np.uint32))


def test_near_ties_duplicates_and_specials_bit_exact(oracle, algo):
    M, k, dsub = 6, 256, 10
    q = random_codebook(M, k, dsub, 5)
    q[:, 17] = q[:, 3]      # exact duplicate centroids: first index must win
    q[:, 200] = q[:, 100]
    x = np.concatenate([near_tie_rows(q, 20_000, 6), np.zeros((3, M * dsub), F), -np.zeros((2, M * dsub), F),
                        np.full((2, M * dsub), 1e-41, F),               # denormals
                        np.full((2, M * dsub), 3.0, F)])               # all-equal rows
    codes = rb.Pq(None, q).quantize_batch(x, np.uint8)
    want = oracle.quantize_batch(q, None, x, np.uint8, n_threads=8)
    assert np.array_equal(codes, want), f"{(codes != want).sum()} codes differ"


def test_nan_and_inf_rank_like_ordered_float(oracle, algo):
    M, k, dsub = 2, 32, 8
    q = random_codebook(M, k, dsub, 9)
    q[0, 0, 0] = np.nan
    q[1, 5, 2] = np.inf
    x = normal((500, M * dsub), 10)
    x[7, 3] = np.nan
    x[8, 9] = np.inf
    x[9, :] = np.nan
    x[10, 0] = -np.inf
    x[11, :] = 1e30   # overflowing norms
    codes = rb.Pq(None, q).quantize_batch(x, np.uint8)
    want = oracle.quantize_batch(q, None, x, np.uint8)
    assert np.array_equal(codes, want)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64, np.int32])
def test_code_types_and_truncating_cast(oracle, dtype):
    q = random_codebook(2, 300, 4, 31)
    x = normal((2_000, 8), 32)
    codes = rb.Pq(None, q).quantize_batch(x, dtype)   # k = 300 > u8: the batch path truncates (primitives.rs:100)
    want = oracle.quantize_batch(q, None, x, np.dtype(dtype).type if np.dtype(dtype).kind == "u" else np.uint32)
    mask = np.uint64(2 ** (8 * np.dtype(dtype).itemsize) - 1)
    assert np.array_equal(codes.astype(np.uint64) & mask, want.astype(np.uint64) & mask)


def test_strided_views_host_and_device(oracle, torch_cuda, algo):
    torch = torch_cuda
    M, k, dsub = 5, 64, 6
    q = random_codebook(M, k, dsub, 41)
    big = normal((4_000, 2 * M * dsub + 3), 42)
    pq = rb.Pq(None, q)
    views = {
        "row_pitch": big[:, :M * dsub],                      # unit column stride, padded rows
        "every_other_row": big[::2, 3:3 + M * dsub],
        "fortran": np.asfortranarray(big[:, :M * dsub]),     # column-major: sequential-dot norms in the reference
        "col_step": big[:, 0:2 * M * dsub:2],
    }
    for name, v in views.items():
        want = oracle.quantize_batch(q, None, v, np.uint8)
        assert np.array_equal(pq.quantize_batch(v, np.uint8), want), name
        t = torch.from_numpy(big).cuda()
        if name == "row_pitch":
            tv = t[:, :M * dsub]
        elif name == "every_other_row":
            tv = t[::2, 3:3 + M * dsub]
        elif name == "fortran":
            tv = t[:, :M * dsub].t().contiguous().t()
        else:
            tv = t[:, 0:2 * M * dsub:2]
        assert np.array_equal(pq.quantize_batch(tv, np.uint8).cpu().numpy(), want), name + " (device)"
    # strided outputs: codes into a column-sliced buffer, reconstructions into a padded buffer
    out = np.zeros((4_000, 2 * M), np.uint8)
    pq.quantize_batch_into(views["row_pitch"], out[:, ::2])
    want = oracle.quantize_batch(q, None, views["row_pitch"], np.uint8)
    assert np.array_equal(out[:, ::2], want) and not out[:, 1::2].any()
    rec = np.zeros((4_000, M * dsub + 5), F)
    pq.reconstruct_batch_into(out[:, ::2], rec[:, 2:2 + M * dsub])
    assert np.array_equal(rec[:, 2:2 + M * dsub], oracle.reconstruct_batch(q, None, want))
    assert not rec[:, :2].any() and not rec[:, 2 + M * dsub:].any()
    rec_t = np.zeros((M * dsub, 4_000), F)
    pq.reconstruct_batch_into(want, rec_t.T)
    assert np.array_equal(rec_t.T, oracle.reconstruct_batch(q, None, want))


def test_empty_batch():
    pq = rb.Pq(None, random_codebook(3, 16, 4, 1))
    assert pq.quantize_batch(np.zeros((0, 12), F)).shape == (0, 3)
    assert pq.reconstruct_batch(np.zeros((0, 3), np.uint8)).shape == (0, 12)


def test_shape_mismatches_panic_like_the_reference():
    pq = rb.Pq(None, random_codebook(3, 16, 4, 1))
    with pytest.raises(rb.ReductivePanic):  # primitives.rs:74-78
        pq.quantize_batch(np.zeros((5, 11), F))
    with pytest.raises(rb.ReductivePanic):  # primitives.rs:80-87
        pq.quantize_batch_into(np.zeros((5, 12), F), np.zeros((5, 4), np.uint8))
    with pytest.raises(rb.ReductivePanic):  # primitives.rs:25-29
        pq.quantize_vector(np.zeros((11,), F))
    with pytest.raises(rb.ReductivePanic):  # primitives.rs:159-167
        pq.reconstruct_batch_into(np.zeros((5, 3), np.uint8), np.zeros((5, 13), F))
    with pytest.raises(IndexError):  # out-of-range code: ndarray index panic in the reference
        pq.reconstruct_batch(np.full((5, 3), 16, np.uint8))
    with pytest.raises(IndexError):
        pq.reconstruct(np.array([0, 99, 0], np.uint8))


def test_vector_paths_bit_exact(oracle):
    M, k, dsub = 4, 128, 9
    q = random_codebook(M, k, dsub, 51)
    x = normal((64, M * dsub), 52)
    r = orthonormal(M * dsub, 53)
    for proj in (None, r):
        pq = rb.Pq(proj, q)
        for i in range(64):
            want = oracle.quantize_vector(q, proj, x[i], np.uint8)
            assert np.array_equal(pq.quantize_vector(x[i], np.uint8), want)
            assert np.array_equal(pq.reconstruct(want).view(np.uint32), oracle.reconstruct(q, proj, want).view(np.uint32))
    xs = normal((2 * M * dsub,), 54)[::2]   # strided vector view
    assert np.array_equal(rb.Pq(None, q).quantize_vector(xs, np.uint8), oracle.quantize_vector(q, None, xs, np.uint8))


@pytest.fixture(params=["exact", "auto"])
def palgo(request):
    rb.set_project_algo(rb.PROJECT_EXACT if request.param == "exact" else rb.PROJECT_AUTO)
    yield request.param
    rb.set_project_algo(rb.PROJECT_AUTO)


@pytest.mark.parametrize("n,M,k,dsub", [(6_000, 10, 256, 30), (6_000, 30, 256, 10), (1_000, 4, 32, 80),
                                        (3_000, 96, 256, 8), (5_000, 16, 128, 8)])
def test_projected_encode_decode(oracle, algo, palgo, n, M, k, dsub):
    """Opq / GaussianOpq use: x.R before the argmin, R^T after the gather (pq.rs:276, 323-326).  d = 300/320 > kc
    exercises the 256-block split of the reference GEMM.  Codes are bit-exact under every rotation kernel; the
    rotated reconstruction is bit-exact with the FP32 GEMM and within north_star's 1e-5 with the tensor-core one."""
    d = M * dsub
    q, r, x = random_codebook(M, k, dsub, 61), orthonormal(d, 62), normal((n, d), 63)
    pq = rb.Pq(r, q)
    codes = pq.quantize_batch(x, np.uint8)
    want = oracle.quantize_batch(q, r, x, np.uint8, n_threads=8)
    assert np.array_equal(codes, want)
    rec = pq.reconstruct_batch(codes)
    want_rec = oracle.reconstruct_batch(q, r, want, n_threads=8)
    if palgo == "exact" or n < 1024:
        assert np.array_equal(rec.view(np.uint32), want_rec.view(np.uint32))
    assert np.abs(rec - want_rec).max() <= 1e-5 * np.abs(want_rec).max()


@pytest.mark.parametrize("M,k,dsub,xscale", [(30, 256, 10, 1.0), (16, 256, 8, 1e-4), (12, 200, 16, 300.0)])
def test_rotated_tensor_encode_adversarial(oracle, M, k, dsub, xscale):
    """Tensor rotation + tensor encode on inputs built to break the certificate: rows that rotate onto (or a few ulp
    off) centroid bisectors and onto exact centroids, NaN / Inf / huge / tiny / zero rows, a row far above the sampled
    operand scale, duplicate centroids, a non-orthonormal R.  Codes must equal the oracle's bit for bit."""
    d = M * dsub
    q = random_codebook(M, k, dsub, 71)
    q[:, 9] = q[:, 4]
    rng = np.random.default_rng(72)
    r = (orthonormal(d, 73) * rng.uniform(0.7, 1.5, size=(1, d))).astype(F)
    y = near_tie_rows(q, 6_000, 74).astype(np.float64)      # wanted rotated rows ...
    x_tie = np.linalg.solve(r.astype(np.float64).T, y.T).T.astype(F)   # ... and rows that rotate (almost) onto them
    x = np.concatenate([x_tie, (normal((4_000, d), 75) * xscale).astype(F)])
    x[10, 3] = np.nan
    x[11, 7] = np.inf
    x[12, :] = -np.inf
    x[13, :] = 1e30
    x[14, :] = 1e-30
    x[15, :] = 0.0
    x[16, :] = 1e-41
    x[17, :] *= 1e6                                          # far above the sampled operand scale
    x[18, 5] = 3e38
    rb.set_project_algo(rb.PROJECT_TENSOR)
    rb.set_encode_algo(rb.ENCODE_AUTO)
    try:
        codes = rb.Pq(r, q).quantize_batch(x, np.uint8)
    finally:
        rb.set_project_algo(rb.PROJECT_AUTO)
    want = oracle.quantize_batch(q, r, x, np.uint8, n_threads=8)
    assert np.array_equal(codes, want), f"{(codes != want).sum()} codes differ in rows {np.unique(np.nonzero(codes != want)[0])[:10]}"


def test_tensor_rotation_on_padded_views(oracle):
    """Device views with a row pitch larger than d on both sides of the tensor rotation (TMA tensor maps with a
    pitch, the exact re-rotation reading the padded rows): codes equal to the oracle's, reconstruction within 1e-5."""
    import torch

    n, M, k, dsub = 5_000, 30, 256, 10
    d = M * dsub
    q, r, x = random_codebook(M, k, dsub, 91), orthonormal(d, 92), normal((n, d), 93)
    pq = rb.Pq(r, q)
    xp = torch.zeros((n, d + 12), device="cuda")
    xp[:, :d] = torch.from_numpy(x).cuda()
    codes = torch.empty((n, M), dtype=torch.uint8, device="cuda")
    recp = torch.full((n, d + 20), 7.0, device="cuda")
    rb.set_project_algo(rb.PROJECT_TENSOR)
    try:
        pq.quantize_batch_into(xp[:, :d], codes)
        pq.reconstruct_batch_into(codes, recp[:, :d])
    finally:
        rb.set_project_algo(rb.PROJECT_AUTO)
    want = oracle.quantize_batch(q, r, x, np.uint8, n_threads=8)
    assert np.array_equal(codes.cpu().numpy(), want)
    want_rec = oracle.reconstruct_batch(q, r, want, n_threads=8)
    rec = recp[:, :d].cpu().numpy()
    assert np.abs(rec - want_rec).max() <= 1e-5 * np.abs(want_rec).max()
    assert bool((recp[:, d:] == 7.0).all())      # nothing written past the view


@pytest.mark.parametrize("n,M,dsub", [(400_000, 16, 8), (150_000, 30, 10)])
def test_tensor_rotation_many_units_per_cta(n, M, dsub):
    """Long runs of (tile, column group) units per CTA: the operand rings wrap many times (a converter set that skipped
    barrier phases once read a stage early here).  Tensor rotation against the exact GPU GEMM on the same codes."""
    import torch

    k, d = 256, M * dsub
    q = random_codebook(M, k, dsub, 81)
    pq = rb.Pq(orthonormal(d, 82), q)
    g = torch.Generator(device="cuda")
    g.manual_seed(83)
    codes = torch.randint(0, k, (n, M), generator=g, device="cuda", dtype=torch.uint8)
    rec_e = torch.empty((n, d), device="cuda")
    rec_t = torch.empty((n, d), device="cuda")
    try:
        rb.set_project_algo(rb.PROJECT_EXACT)
        pq.reconstruct_batch_into(codes, rec_e)
        rb.set_project_algo(rb.PROJECT_TENSOR)
        for _ in range(3):
            pq.reconstruct_batch_into(codes, rec_t)
    finally:
        rb.set_project_algo(rb.PROJECT_AUTO)
    torch.cuda.synchronize()
    assert float((rec_e - rec_t).abs().max()) <= 1e-5 * float(rec_e.abs().max())


@pytest.mark.parametrize("d,scale", [(300, 1.0), (768, 1e-3), (128, 37.0), (36, 1e4), (260, 1.0)])
def test_tensor_rotation_decode_error(d, scale):
    """The tcgen05 rotation against float64: error inside north_star's 1e-5 of the largest output, for codebooks of very
    different magnitudes (the operand scale is a power of two derived from |centroid|max) and a non-orthonormal R."""
    dsub = 4
    M, k, n = d // dsub, 64, 5_000
    rng = np.random.default_rng(d)
    q = (rng.normal(size=(M, k, dsub)) * scale).astype(F)
    q[0, 0, 0] = 0.0
    r = (orthonormal(d, 7) * rng.uniform(0.5, 2.0, size=(1, d))).astype(F)
    codes = rng.integers(0, k, size=(n, M)).astype(np.uint8)
    pq = rb.Pq(r, q)
    rb.set_project_algo(rb.PROJECT_TENSOR)
    try:
        rec = pq.reconstruct_batch(codes)
    finally:
        rb.set_project_algo(rb.PROJECT_AUTO)
    flat = q[np.arange(M)[None, :], codes.astype(np.int64)].reshape(n, d).astype(np.float64)
    want = flat @ r.astype(np.float64).T
    assert np.abs(rec - want).max() <= 1e-5 * np.abs(want).max()


# ---- k-means / training ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n,M,bits,dsub,iters", [(10_000, 10, 8, 30, 10), (20_000, 12, 8, 8, 6), (3_000, 5, 4, 6, 8)])
def test_train_pq_bit_exact_from_identical_initial_centroids(oracle, algo, n, M, bits, dsub, iters):
    """Single GPU, reference-order centroid update: assignments are bit-exact and the per-cluster sums are added in
    the reference's row order, so the trained codebook is IDENTICAL to the oracle's (north_star asks <= 1e-4)."""
    x = normal((n, M * dsub), 71)
    init = rows_as_initial_centroids(x, M, 1 << bits, 72)
    pq, loss = rb.Pq.train_pq_using(M, bits, iters, 1, x, None, initial_centroids=init, return_loss=True)
    want_q, want_loss = oracle.train_pq(x, M, bits, iters, 1, init, n_threads=8)
    got = pq.subquantizers()
    assert np.array_equal(got.view(np.uint32), want_q.view(np.uint32)), \
        f"max rel err {np.abs(got - want_q).max() / np.abs(want_q).max()}"
    assert np.allclose(loss, want_loss, rtol=1e-3)


def test_train_pq_atomic_update_within_tolerance(oracle):
    """The unordered (shared-memory atomics) update differs from the reference by summation order only; with
    ~1000 rows per cluster the trained codebook stays within north_star's 1e-4 relative."""
    n, M, bits, dsub, iters = 60_000, 6, 6, 8, 5
    x = normal((n, M * dsub), 73)
    init = rows_as_initial_centroids(x, M, 1 << bits, 74)
    rb.set_kmeans_update(False)
    try:
        pq = rb.Pq.train_pq_using(M, bits, iters, 1, x, None, initial_centroids=init)
    finally:
        rb.set_kmeans_update(True)
    want_q, _ = oracle.train_pq(x, M, bits, iters, 1, init, n_threads=8)
    got = pq.subquantizers()
    for m in range(M):
        rel = np.linalg.norm(got[m] - want_q[m]) / np.linalg.norm(want_q[m])
        assert rel <= 1e-4, f"subquantizer {m}: relative error {rel}"


def test_single_kmeans_iteration_is_tight(oracle, torch_cuda):
    """One Lloyd step has no chaotic amplification: identical assignments, sums equal to summation-order error."""
    torch = torch_cuda
    x = normal((50_000, 8), 81)
    init = rows_as_initial_centroids(x, 1, 256, 82)[0, 0]
    want_c, want_loss = oracle.kmeans_iteration(x, init)
    cen = torch.from_numpy(init.copy()).cuda()
    loss = rb.kmeans_iteration(torch.from_numpy(x).cuda(), cen)
    assert np.array_equal(cen.cpu().numpy().view(np.uint32), want_c.view(np.uint32))
    # the reference sums 400k squared errors sequentially in f32 (kmeans.rs:357): ITS value drifts ~1e-4 from the
    # exactly-rounded one this path computes in f64
    assert abs(loss - want_loss) <= 1e-3 * want_loss


def test_empty_clusters_stay_zero(oracle, torch_cuda):  # kmeans.rs:181,194
    torch = torch_cuda
    x = normal((400, 4), 91)
    cen0 = np.concatenate([x[:6], np.full((2, 4), 1e3, F)])   # two centroids nobody is assigned to
    want_c, _ = oracle.kmeans_iteration(x, cen0)
    cen = torch.from_numpy(cen0.copy()).cuda()
    rb.kmeans_iteration(torch.from_numpy(x).cuda(), cen)
    got = cen.cpu().numpy()
    assert not got[6:].any() and not want_c[6:].any()
    assert np.array_equal(got.view(np.uint32), want_c.view(np.uint32))


def test_multi_attempt_training_picks_lowest_loss(oracle):
    x = normal((2_000, 8), 95)
    init = rows_as_initial_centroids(x, 2, 8, 96, n_attempts=3)
    pq, loss = rb.Pq.train_pq_using(2, 3, 4, 3, x, None, initial_centroids=init, return_loss=True)
    want_q, want_loss = oracle.train_pq(x, 2, 3, 4, 3, init)
    assert np.allclose(loss, want_loss, rtol=1e-3)
    assert np.linalg.norm(pq.subquantizers() - want_q) / np.linalg.norm(want_q) <= 1e-4


def test_train_validation_errors():
    x = normal((256, 20), 1)
    with pytest.raises(rb.IncorrectNSubquantizerBits):
        rb.Pq.train_pq_using(10, 9, 10, 1, x, np.random.default_rng(0))
    with pytest.raises(rb.IncorrectNumberSubquantizers):
        rb.Pq.train_pq_using(3, 7, 10, 1, x, np.random.default_rng(0))
    with pytest.raises(rb.ReductivePanic):  # k == n: RandomInstanceCentroids asserts k < n (kmeans.rs:62-67)
        rb.Pq.train_pq_using(10, 8, 10, 1, x, np.random.default_rng(0))


# ---- the reference's statistical end-to-end tests ------------------------------------------------------
def _avg_euclidean_loss(x, pq):  # pq.rs:365-376
    rec = pq.reconstruct_batch(pq.quantize_batch(x, np.uint8))
    return float(np.mean(np.sqrt(((x - rec) ** 2).sum(1))))


def test_quantize_with_pq():  # pq.rs:431-440
    rng = np.random.default_rng(42)
    x = rng.random((256, 20), F)
    assert _avg_euclidean_loss(x, rb.Pq.train_pq_using(10, 7, 10, 1, x, rng)) < 0.08


def test_quantize_with_opq():  # opq.rs:330-339
    rng = np.random.default_rng(42)
    x = rng.random((256, 20), F)
    pq = rb.Opq.train_pq_using(10, 7, 10, 1, x, rng)
    assert pq.projection() is not None and _avg_euclidean_loss(x, pq) < 0.1


def test_quantize_with_gaussian_opq():  # gaussian_opq.rs:99-108
    rng = np.random.default_rng(42)
    x = rng.random((256, 20), F)
    pq = rb.GaussianOpq.train_pq_using(10, 7, 10, 1, x, rng)
    assert pq.projection() is not None and _avg_euclidean_loss(x, pq) < 0.12


def test_k_means_3():  # kmeans.rs:459-479
    rng = np.random.default_rng(3)
    centers = np.array([[0., 0.], [1., 0.], [1., 1.]], F)
    x = np.concatenate([c + rng.normal(0, 0.01, (11, 2)).astype(F) for c in centers]).astype(F)

    class OnePerBlob:
        def initial_centroids(self, data, k):
            return data[[0, 11, 22]].clone()

    cen, _ = rb.KMeans.k_means(x, 3, OnePerBlob(), rb.NIterationsCondition(10))
    assert sorted(map(tuple, np.rint(cen).astype(int).tolist())) == [(0, 0), (1, 0), (1, 1)]


def test_chained_accumulate_equals_one_pass(oracle, torch_cuda):
    """rb_kmeans_assign + rb_kmeans_accumulate with packed_before: the chains of a second block of rows continue the
    first block's sums, so a split pass is bit-identical to one pass and to the oracle (kmeans.rs:185-189).  Sizes
    straddle the sort chunk (32 768 rows at d = 40) so several chunks and a ragged tail are covered."""
    torch = torch_cuda
    from reductive_b200.dist import cuda_accumulate, cuda_assign, cuda_finalize, cuda_local_step

    n, M, k, dsub = 70_001, 5, 64, 8
    x = normal((n, M * dsub), 101)
    init = rows_as_initial_centroids(x, M, k, 102)[0]
    xd = torch.from_numpy(x).cuda()
    cen = torch.from_numpy(init.copy()).cuda()
    plen = M * k * dsub + M * k + M
    one = torch.empty((plen,), dtype=torch.float32, device="cuda")
    cuda_local_step(xd, cen, one)
    n0 = M * k * dsub + M * k  # sums and counts (the sum of squared norms is accumulated with float atomics)
    for cut in (1, 33_000, 65_536, n - 1):
        a, b = xd[:cut], xd[cut:]
        p0 = torch.empty_like(one)
        p1 = torch.empty_like(one)
        cuda_accumulate(a, cen, cuda_assign(a, cen), None, p0)
        cuda_accumulate(b, cen, cuda_assign(b, cen), p0, p1)
        assert torch.equal(p1[:n0].view(torch.int32), one[:n0].view(torch.int32)), cut
        assert torch.allclose(p1[n0:], one[n0:], rtol=1e-5)
    # a slice of the subquantizers [1, 4) accumulated on its own (what the chained multi-GPU mode relays) equals the
    # same slice of the full pass
    from reductive_b200.dist import slice_len
    m0, m1 = 1, 4
    ps = torch.empty((slice_len(m0, m1, k, dsub),), dtype=torch.float32, device="cuda")
    codes_all = cuda_assign(xd, cen)
    cuda_accumulate(xd, cen, codes_all, None, ps, m0, m1)
    ns = (m1 - m0) * k * dsub
    assert torch.equal(ps[:ns].view(torch.int32), one[m0 * k * dsub:m1 * k * dsub].view(torch.int32))
    assert torch.equal(ps[ns:ns + (m1 - m0) * k], one[M * k * dsub + m0 * k:M * k * dsub + m1 * k])
    # and the finalized centroids are the oracle's
    loss = torch.zeros((M,), device="cuda")
    cuda_finalize(one, n, cen, loss)
    want = np.stack([oracle.kmeans_iteration(np.ascontiguousarray(x[:, m * dsub:(m + 1) * dsub]), init[m])[0]
                     for m in range(M)])
    assert np.array_equal(cen.cpu().numpy().view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("n,M,k,dsub", [
    (100_003, 30, 256, 10),   # C2 geometry: three column groups in a cluster, two lookups per 16-byte piece, ragged n*M
    (70_001, 96, 256, 8),     # C3 geometry: eight line-aligned groups, small code tiles
    (50_000, 16, 200, 8),     # k < 256: range-checked codes
    (33_333, 50, 64, 6),      # dsub % 4 == 2 with several groups
    (300, 12, 256, 4),        # one tile, fewer rows than a tile
])
def test_tiled_gather_bit_exact_and_range_checked(oracle, torch_cuda, n, M, k, dsub):
    """gather_tile_kernel (TMA-staged code ring, fixed column per thread, cluster lockstep) against the oracle's
    copy loop (primitives.rs:110-173), into a dense and into a row-padded output; a code >= k is reported."""
    torch = torch_cuda
    q = random_codebook(M, k, dsub, 300 + n)
    codes = np.random.default_rng(n).integers(0, k, size=(n, M)).astype(np.uint8)
    want = oracle.reconstruct_batch(q, None, codes, n_threads=8)
    pq = rb.Pq(None, q)
    cd = torch.from_numpy(codes).cuda()
    rec = pq.reconstruct_batch(cd)
    assert np.array_equal(rec.cpu().numpy().view(np.uint32), want.view(np.uint32))
    padded = torch.full((n, M * dsub + 8), -7.0, device="cuda")
    pq.reconstruct_batch_into(cd, padded[:, :M * dsub])
    got = padded.cpu().numpy()
    assert np.array_equal(got[:, :M * dsub].view(np.uint32), want.view(np.uint32))
    assert (got[:, M * dsub:] == -7.0).all()  # nothing written past the row
    if k < 256:
        bad = cd.clone()
        bad[n // 2, M - 1] = k
        with pytest.raises(IndexError):  # the reference's ndarray index panic
            pq.reconstruct_batch(bad)


def test_padded_codebook_large_norm_rows_and_ties(oracle, algo):
    """64 < k < 256 runs the tensor kernel with padding columns of a fixed large score: rows whose norm could bring a
    real score near it must be decided exactly, near ties as usual."""
    M, k, dsub = 4, 128, 10
    q = random_codebook(M, k, dsub, 77)
    x = np.concatenate([near_tie_rows(q, 4_000, 78), normal((2_000, M * dsub), 79),
                        normal((64, M * dsub), 80) * F(1e3), normal((64, M * dsub), 81) * F(1e5),
                        normal((64, M * dsub), 82) * F(1e-6)])
    codes = rb.Pq(None, q).quantize_batch(x, np.uint8)
    want = oracle.quantize_batch(q, None, x, np.uint8, n_threads=8)
    assert codes.max() < k
    assert np.array_equal(codes, want), f"{(codes != want).sum()} codes differ"


@pytest.mark.parametrize("dsub", [2, 4, 6, 8, 10, 12, 16, 20, 24, 30, 32])
@pytest.mark.parametrize("scale", [1e-4, 1.0, 3e3])
def test_tensor_margin_holds_across_scales_and_widths(oracle, dsub, scale):
    """The tensor pass may only decide what the reference decides the same way: near-tie rows (bisector +- ulps),
    random rows and rows far outside the codebook, at three magnitudes and every instantiated subvector width,
    with a codebook whose subquantizers differ in scale by 64x."""
    M, k = 4, 256
    q = random_codebook(M, k, dsub, 1000 + dsub)
    q *= np.array([1.0, 8.0, 0.125, 1.0], F)[:, None, None]
    q = (q * F(scale)).astype(F)
    x = np.concatenate([near_tie_rows(q, 6_000, 2000 + dsub), normal((3_000, M * dsub), 3000 + dsub) * F(scale),
                        normal((500, M * dsub), 4000 + dsub) * F(30.0 * scale)]).astype(F)
    rb.set_encode_algo(rb.ENCODE_AUTO)
    codes = rb.Pq(None, q).quantize_batch(x, np.uint8)
    want = oracle.quantize_batch(q, None, x, np.uint8, n_threads=8)
    assert np.array_equal(codes, want), f"{(codes != want).sum()} codes differ"


@pytest.mark.parametrize("dsub", [8, 10])
@pytest.mark.parametrize("spread", [(1.0, 1e3, 1e-3, 1.0), (1e4, 1.0, 1e-4, 3e2)])
def test_tensor_margin_with_wide_subquantizer_scale_spread(oracle, dsub, spread):
    """with ONE operand scale for all subquantizers, a subquantizer 1e3-1e4x smaller than the largest fell
    into FP16 subnormals and the certificate could certify a wrong winner.  The scale is per subquantizer now; near-tie
    rows in every subquantizer of codebooks with 1e3x and 1e8x spreads must come out bit-exact."""
    M, k = 4, 256
    q = random_codebook(M, k, dsub, 5000 + dsub)
    q = (q * np.array(spread, F)[:, None, None]).astype(F)
    rows = near_tie_rows(q, 8_000, 6000 + dsub)
    rnd = normal((3_000, M * dsub), 7000 + dsub) * np.repeat(np.array(spread, F), dsub)[None, :]
    x = np.concatenate([rows, rnd.astype(F)]).astype(F)
    rb.set_encode_algo(rb.ENCODE_AUTO)
    codes = rb.Pq(None, q).quantize_batch(x, np.uint8)
    want = oracle.quantize_batch(q, None, x, np.uint8, n_threads=8)
    assert np.array_equal(codes, want), f"{(codes != want).sum()} codes differ"


def test_pageable_and_pinned_host_buffers_give_the_same_codes(oracle, torch_cuda):
    """The host-memory pipeline stages pageable memory through pinned buffers of its own (several chunks, both
    slots, strided rows); pinned memory is copied from directly."""
    torch = torch_cuda
    n, M, k, dsub = 200_000, 10, 256, 10  # 240 MB: four 64 MB chunks
    q = random_codebook(M, k, dsub, 11)
    pq = rb.Pq(None, q)
    x = normal((n, M * dsub + 4), 12)[:, : M * dsub]  # row pitch > d: strided rows
    codes = pq.quantize_batch(x, np.uint8)
    xp = torch.empty((n, M * dsub), dtype=torch.float32, pin_memory=True)
    xp.copy_(torch.from_numpy(np.ascontiguousarray(x)))
    cp = torch.empty((n, M), dtype=torch.uint8, pin_memory=True)
    pq.quantize_batch_into(xp.numpy(), cp.numpy())
    assert np.array_equal(codes, cp.numpy())
    sample = np.random.default_rng(13).choice(n, 3000, replace=False)
    assert np.array_equal(codes[sample], oracle.quantize_batch(q, None, np.ascontiguousarray(x[sample]), np.uint8))
    rec = pq.reconstruct_batch(codes)
    assert np.array_equal(rec[sample], oracle.reconstruct_batch(q, None, codes[sample]))


def test_covariance_matches_the_oracle(oracle, torch_cuda):  # linalg.rs:23-44, KAT linalg.rs:254-262
    import ctypes as C  # noqa: F401

    torch = torch_cuda
    from reductive_b200._cabi import check, lib

    def cov(x):
        xd = torch.from_numpy(x).cuda()
        out = torch.empty((x.shape[1], x.shape[1]), device="cuda")
        check(lib.rb_covariance(xd.data_ptr(), x.shape[0], x.shape[1], xd.stride(0), out.data_ptr(), None))
        return out.cpu().numpy()

    assert np.array_equal(cov(np.array([[0., 2.], [1., 1.], [2., 0.]], F)), np.array([[1., -1.], [-1., 1.]], F))
    x = normal((20_000, 96), 71) * np.linspace(0.5, 3, 96, dtype=F) + np.linspace(-2, 2, 96, dtype=F)
    got, want = cov(x), oracle.covariance(x)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    assert np.array_equal(got, got.T) or np.abs(got - got.T).max() <= 1e-6 * np.abs(got).max()


@pytest.mark.parametrize("n,d,pad", [(50_000, 300, 0), (5_000, 64, 0), (20_001, 512, 0), (9_000, 137, 3), (4_096, 390, 0),
                                     (300_000, 96, 0), (7_777, 16, 5)])
def test_tensor_core_gram_matches_float64(torch_cuda, n, d, pad):
    """gram_tc.cu (rb_set_gram_algo(2)): the centred Gram matrix over the rows on tcgen05 -- two-limb BF16 operands,
    FP32 accumulation in 2048-row segments -- against the float64 product, and beside the FP32 CUDA-core kernel.  Both
    are held to 1e-5 of the largest element (the tolerance of rb_covariance / X^T.Y^).  Widths cover one and several
    128-column tiles, one and two MMA column blocks, widths that are no multiple of 16, and padded rows."""
    torch = torch_cuda
    from reductive_b200._cabi import GRAM_AUTO, GRAM_CUDA_CORES, GRAM_TENSOR, check, lib, set_gram_algo

    rng = np.random.default_rng(n + d)
    a = (rng.normal(size=(n, d + pad)) * np.linspace(0.2, 4, d + pad) + np.linspace(-3, 3, d + pad)).astype(F)
    a[:, 1] += 0.5 * a[:, 0]  # some real off-diagonal mass
    ad = torch.from_numpy(a).cuda()
    c = a[:, :d].astype(np.float64) - a[:, :d].mean(0, dtype=np.float64)
    want = c.T @ c / (n - 1)

    def cov(algo):
        set_gram_algo(algo)
        try:
            out = torch.empty((d, d), device="cuda")
            check(lib.rb_covariance(ad.data_ptr(), n, d, ad.stride(0), out.data_ptr(), None))
            return out.cpu().numpy()
        finally:
            set_gram_algo(GRAM_AUTO)

    got_tc, got_cc = cov(GRAM_TENSOR), cov(GRAM_CUDA_CORES)
    scale = np.abs(want).max()
    assert np.abs(got_tc - want).max() <= 1e-5 * scale, np.abs(got_tc - want).max() / scale
    assert np.abs(got_cc - want).max() <= 1e-5 * scale
    assert np.abs(got_tc - got_tc.T).max() <= 4e-6 * scale


def test_tensor_core_gram_in_several_row_passes(torch_cuda, monkeypatch):
    """Large inputs go through gram_tc.cu in row passes that share one partial buffer (RB_GRAM_PASS_ROWS forces the
    passes small here): the same tolerance, and the ragged last pass counts."""
    torch = torch_cuda
    from reductive_b200._cabi import GRAM_AUTO, GRAM_TENSOR, check, lib, set_gram_algo

    n, d = 21_003, 200
    a = (normal((n, d), 123) * np.linspace(0.5, 2, d, dtype=F) + F(0.3)).astype(F)
    ad = torch.from_numpy(a).cuda()
    c = a.astype(np.float64) - a.mean(0, dtype=np.float64)
    want = c.T @ c / (n - 1)
    monkeypatch.setenv("RB_GRAM_PASS_ROWS", "8192")
    set_gram_algo(GRAM_TENSOR)
    try:
        out = torch.empty((d, d), device="cuda")
        check(lib.rb_covariance(ad.data_ptr(), n, d, ad.stride(0), out.data_ptr(), None))
    finally:
        set_gram_algo(GRAM_AUTO)
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-5 * np.abs(want).max()


@pytest.mark.parametrize("n,M,k,dsub", [(6_000, 4, 16, 6), (20_000, 8, 256, 8)])
def test_opq_train_iteration_matches_the_oracle(oracle, torch_cuda, n, M, k, dsub):
    """Opq::train_iteration (opq.rs:161-189) from identical projection and centroids: the k-means step is bit-identical
    (same rx: the projection kernel keeps the reference's summation order), X^T.Y^ and the new rotation agree within
    float summation-order tolerance (1e-5 / 1e-4 relative)."""
    torch = torch_cuda
    from reductive_b200._cabi import check, lib

    d = M * dsub
    x = normal((n, d), 81) * np.linspace(0.3, 2, d, dtype=F)
    r0 = oracle.create_projection_matrix(x, M)
    rx = oracle.sgemm(x, r0)
    c0 = rows_as_initial_centroids(rx, M, k, 82)[0]
    want_r, want_c, want_xty = oracle.opq_train_iteration(r0, c0, x)
    xd, rd, cd = torch.from_numpy(x).cuda(), torch.from_numpy(r0).cuda(), torch.from_numpy(c0).cuda()
    xty = torch.empty((d, d), device="cuda")
    check(lib.rb_opq_train_iteration(xd.data_ptr(), n, d, xd.stride(0), rd.data_ptr(), cd.data_ptr(), M, k, xty.data_ptr(), None))
    assert np.array_equal(cd.cpu().numpy(), want_c), "k-means step differs from the oracle"
    got_xty = xty.cpu().numpy()
    assert np.abs(got_xty - want_xty).max() <= 1e-5 * np.abs(want_xty).max()
    u, _, vt = np.linalg.svd(got_xty, full_matrices=True)
    assert np.abs(u @ vt - want_r).max() <= 1e-4


def test_streaming_update_gives_the_same_bits(oracle):
    """rb_set_kmeans_update(3): the training loop streams subquantizer-major slabs (kmeans.cu ordered_stream_kernel);
    the sums are the same sequential f32 chains, so the trained codebook is bit-identical (uneven tile, small k)."""
    x = normal((7_013, 48), 91)
    for M, bits in ((6, 8), (4, 5)):
        init = rows_as_initial_centroids(x, M, 1 << bits, 92)
        want, _ = oracle.train_pq(x, M, bits, 4, 1, init, n_threads=4)
        rb.set_kmeans_update(3)
        try:
            got = rb.Pq.train_pq_using(M, bits, 4, 1, x, None, initial_centroids=init).subquantizers()
        finally:
            rb.set_kmeans_update(1)
        assert np.array_equal(got.view(np.int32), want.view(np.int32))
