"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header declares, the
host-only entry points agree with the oracle, and compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import reductive_b200 as rb
from reductive_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "reductive_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    declared = _header_functions()
    assert declared, "no declarations parsed from the header"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(_cabi.lib, name), f"{name} is declared in include/reductive_b200.h but not exported"
    assert _cabi.lib.rb_abi_version() == 3


def test_check_quantizer_invariants_matches_oracle(oracle):
    cases = [(10, 7, 10, 1, 256, 20), (0, 7, 10, 1, 256, 20), (21, 7, 10, 1, 256, 20), (10, 0, 10, 1, 256, 20),
             (10, 9, 10, 1, 256, 20), (10, 8, 10, 1, 256, 20), (10, 8, 10, 1, 255, 20), (3, 7, 10, 1, 256, 20),
             (10, 7, 0, 1, 256, 20), (10, 7, 10, 0, 256, 20), (10, 1, 10, 1, 0, 20), (30, 8, 1, 1, 2_000_000, 300),
             (96, 8, 25, 1, 1_000_000, 768), (7, 3, 1, 1, 9, 7)]
    for c in cases:
        want, want_detail = oracle.check_quantizer_invariants(*c)
        detail = C.c_uint64(0)
        got = _cabi.lib.rb_check_quantizer_invariants(*c, C.byref(detail))
        assert got == want, c
        if want in (3, 5):
            assert detail.value == want_detail, c


def test_error_classes_mirror_reductive_error():
    with pytest.raises(rb.NSubquantizersOutsideRange):
        rb.check_quantizer_invariants(0, 7, 10, 1, 256, 20)
    with pytest.raises(rb.IncorrectNSubquantizerBits):
        rb.check_quantizer_invariants(10, 9, 10, 1, 256, 20)
    with pytest.raises(rb.IncorrectNumberSubquantizers):
        rb.check_quantizer_invariants(3, 7, 10, 1, 256, 20)
    with pytest.raises(rb.IncorrectNIterations):
        rb.check_quantizer_invariants(10, 7, 0, 1, 256, 20)
    with pytest.raises(rb.IncorrectNAttempts):
        rb.check_quantizer_invariants(10, 7, 10, 0, 256, 20)


def test_packed_len():
    assert _cabi.lib.rb_kmeans_packed_len(96, 256, 8) == 96 * 256 * 8 + 96 * 256 + 96


def test_pq_new_panics_like_the_reference():
    with pytest.raises(rb.ReductivePanic):  # pq.rs:39-42
        rb.Pq(None, np.zeros((0, 4, 3), np.float32))
    with pytest.raises(rb.ReductivePanic):  # pq.rs:46-55
        rb.Pq(np.eye(5, dtype=np.float32), np.zeros((2, 4, 3), np.float32))


def test_bucket_eigenvalues():  # opq.rs:303-311
    assert rb.bucket_eigenvalues(np.array([0.2, 0.6, 0.4, 0.1, 0.3, 0.5]), 3) == [[1, 3], [5, 0], [2, 4]]


def test_bucket_large_eigenvalues():  # opq.rs:313-320
    ev = np.array([11174., 23450., 30835., 1557., 32425., 5154.])
    assert rb.bucket_eigenvalues(ev, 3) == [[4, 3], [2, 5], [1, 0]]


def test_bucket_eigenvalues_uneven():  # opq.rs:322-328 (#[should_panic])
    with pytest.raises(rb.ReductivePanic):
        rb.bucket_eigenvalues(np.array([0.2, 0.6, 0.4, 0.1, 0.3, 0.5]), 4)


def test_random_instance_centroids_are_distinct_rows():
    from reductive_b200.pq import random_instance_centroids

    idx = random_instance_centroids((300, 20), 10, 128, 2, np.random.default_rng(0))
    assert idx.shape == (2, 10, 128)
    for a in range(2):
        for m in range(10):
            assert len(set(idx[a, m].tolist())) == 128 and idx[a, m].min() >= 0 and idx[a, m].max() < 300
    with pytest.raises(rb.ReductivePanic):  # kmeans.rs:62-67
        random_instance_centroids((128, 20), 10, 128, 1, np.random.default_rng(0))


def test_compute_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(rb.NoDeviceError):
        rb.Pq(None, np.ones((2, 4, 3), np.float32))
    with pytest.raises(rb.NoDeviceError):
        rb.Pq.train_pq_using(2, 2, 1, 1, np.ones((16, 4), np.float32), np.random.default_rng(0))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "reductive_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "oracle.h" not in text, f


def test_shard_rows_covers_everything():
    from reductive_b200.dist import shard_rows

    for n in (0, 1, 7, 8, 1000, 2_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_rows(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0 and a0 <= a1


def test_subquantizer_ranges_partition_the_subquantizers():
    """rb_dist_subquantizer_range (host only): the ranks' ranges are contiguous, disjoint, cover [0, M), differ in
    size by at most one, and ranks beyond M own nothing."""
    from reductive_b200.dist import subquantizer_range

    for M in (1, 2, 5, 16, 30, 96, 97):
        for world in (1, 2, 3, 4, 8, 16):
            end = 0
            sizes = []
            for r in range(world):
                a, b = subquantizer_range(M, r, world)
                assert a == end and b >= a
                end = b
                sizes.append(b - a)
            assert end == M and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        subquantizer_range(8, 3, 3)


def test_multi_gpu_entry_points_fail_loudly_without_a_device():
    """No CPU fallback: the communicator needs NCCL + a device, training on a device list needs the devices."""
    import ctypes as C

    import torch

    from reductive_b200._cabi import CudaError, check, lib

    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    ident = (C.c_ubyte * 128)()
    h = C.c_void_p()
    with pytest.raises(CudaError):
        check(lib.rb_comm_create(ident, 0, 1, C.byref(h)))


def test_f64_entry_points_answer_unsupported():
    """Pq<f64> (pq.rs:29-32): typed entry points exist and say UNSUPPORTED instead of down-converting silently."""
    import ctypes as C

    from reductive_b200._cabi import ERR_UNSUPPORTED, last_error, lib

    h = C.c_void_p()
    q = np.zeros((1, 2, 2), np.float64)
    lib.rb_pq_create_f64.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]
    assert lib.rb_pq_create_f64(q.ctypes.data, 1, 2, 2, None, C.byref(h)) == ERR_UNSUPPORTED
    assert "f32 only" in last_error() and not h.value


def test_nccl_is_resolved_to_the_copy_torch_uses():
    """The library dlopens NCCL lazily; if that happened before `import torch` with another libnccl.so.2 than the one
    torch is linked against, the later torch import would fail on a missing symbol.  Fresh interpreter: resolve NCCL
    through the library first, then import torch."""
    import subprocess
    import sys

    code = ("import ctypes, reductive_b200._cabi as c\n"
            "buf = (ctypes.c_ubyte * 128)()\n"
            "assert c.lib.rb_comm_unique_id(buf, 128) == 0, c.last_error()\n"
            "import torch\n"
            "print('ok', torch.__version__)\n")
    env = {k: v for k, v in os.environ.items() if k != "RB_NCCL_LIB"}
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_new_entry_points_validate_their_arguments_without_a_device():
    """Argument errors are reported before any device work (no GPU needed)."""
    import ctypes as C

    from reductive_b200._cabi import lib

    assert lib.rb_set_gram_algo(7) != 0 and lib.rb_set_gram_algo(0) == 0
    h = C.c_void_p()
    assert lib.rb_qstore_create(None, None, 0, 0, None, 0, None, C.byref(h)) != 0 and not h.value
    assert lib.rb_qstore_len(None) == 0 and lib.rb_qstore_has_norms(None) == 0
    assert lib.rb_qstore_dot(None, None, 1, 1, 1, None, 1, 0, None) != 0
    assert lib.rb_qstore_embeddings(None, None, 1, None, 1, 1, 0, None) != 0
    assert lib.rb_kmeans_dist_peer_window(None) == 0
    lib.rb_qstore_destroy(None)  # no-op
