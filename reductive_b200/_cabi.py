"""ctypes binding of include/reductive_b200.h (the C ABI of the CUDA library).

There is no CPU implementation behind this module: if the shared library is missing the import fails, and
every compute call fails with NoDeviceError when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RB_LIB_PATH: development aid for A/B runs of two builds of the same library (never a different backend)
LIB_PATH = os.environ.get("RB_LIB_PATH") or os.path.join(_HERE, "lib", "libreductive_b200.so")

# rb_status
OK = 0
ERR_N_ATTEMPTS = 1
ERR_N_ITERATIONS = 2
ERR_N_SUBQUANTIZER_BITS = 3
ERR_NUMBER_SUBQUANTIZERS = 4
ERR_N_SUBQUANTIZERS_RANGE = 5
ERR_CONSTRUCT_RNG = 6
ERR_SHAPE = 16
ERR_CODE_TYPE = 17
ERR_CODE_RANGE = 18
ERR_INVALID = 19
ERR_K_MEANS_K = 20
ERR_CUDA = 32
ERR_NO_DEVICE = 33
ERR_UNSUPPORTED = 34
ERR_NCCL = 35

MEM_HOST, MEM_DEVICE = 0, 1
ENCODE_AUTO, ENCODE_EXACT, ENCODE_TENSOR = 0, 1, 2
PROJECT_AUTO, PROJECT_EXACT, PROJECT_TENSOR = 0, 1, 2

# every symbol include/reductive_b200.h declares (tests check the library exports each one)
EXPORTED_SYMBOLS = [
    "rb_last_error_message", "rb_abi_version", "rb_kernel_launch_count", "rb_set_encode_algo",
    "rb_set_kmeans_update", "rb_set_project_algo", "rb_release_scratch",
    "rb_pq_create", "rb_pq_create_multi", "rb_pq_n_devices", "rb_set_host_copy_threads", "rb_pq_destroy", "rb_pq_quantized_len", "rb_pq_reconstructed_len",
    "rb_pq_n_quantizer_centroids", "rb_pq_has_projection", "rb_pq_subquantizers", "rb_pq_projection",
    "rb_pq_quantize_batch", "rb_pq_quantize_vector", "rb_pq_reconstruct_batch", "rb_pq_reconstruct",
    "rb_check_quantizer_invariants", "rb_kmeans_packed_len", "rb_kmeans_assign_accumulate",
    "rb_kmeans_assign_accumulate_from", "rb_kmeans_code_pitch", "rb_kmeans_code_width", "rb_kmeans_assign",
    "rb_kmeans_accumulate",
    "rb_kmeans_finalize", "rb_pq_train", "rb_project_rows",
    "rb_dist_subquantizer_range", "rb_comm_unique_id", "rb_comm_create", "rb_comm_destroy", "rb_comm_rank",
    "rb_comm_world", "rb_kmeans_dist_create", "rb_kmeans_dist_iterate", "rb_kmeans_dist_destroy", "rb_kmeans_dist_peer_window", "rb_pq_train_dist",
    "rb_pq_train_multi", "rb_set_gram_algo", "rb_covariance", "rb_opq_train_iteration", "rb_pq_create_f64", "rb_pq_quantize_batch_f64", "rb_pq_reconstruct_batch_f64", "rb_pq_train_f64",
    "rb_qstore_create", "rb_qstore_destroy", "rb_qstore_len", "rb_qstore_has_norms", "rb_qstore_embeddings", "rb_qstore_dot",
]


class ReductiveError(Exception):
    """Mirror of reductive::error::ReductiveError (src/error.rs:6-41)."""

    status = -1


class IncorrectNAttempts(ReductiveError):
    status = ERR_N_ATTEMPTS


class IncorrectNIterations(ReductiveError):
    status = ERR_N_ITERATIONS


class IncorrectNSubquantizerBits(ReductiveError):
    status = ERR_N_SUBQUANTIZER_BITS


class IncorrectNumberSubquantizers(ReductiveError):
    status = ERR_NUMBER_SUBQUANTIZERS


class NSubquantizersOutsideRange(ReductiveError):
    status = ERR_N_SUBQUANTIZERS_RANGE


class ConstructRng(ReductiveError):
    status = ERR_CONSTRUCT_RNG


class ReductivePanic(AssertionError):
    """The reference panics (assert!) in this situation; the C ABI returns a status instead."""


class CudaError(RuntimeError):
    pass


class NoDeviceError(CudaError):
    pass


_ERR_CLASSES = {c.status: c for c in (IncorrectNAttempts, IncorrectNIterations, IncorrectNSubquantizerBits,
                                      IncorrectNumberSubquantizers, NSubquantizersOutsideRange, ConstructRng)}


def _point_at_bundled_nccl() -> None:
    """The library resolves NCCL with dlopen when a communicator spans several devices.  This interpreter's torch
    carries its own libnccl.so.2 (nvidia/nccl/lib) and the dynamic loader keeps whichever copy of that soname was
    loaded first, so both must use the same file: name it in RB_NCCL_LIB unless the caller already chose one."""
    if os.environ.get("RB_NCCL_LIB"):
        return
    import importlib.util

    spec = importlib.util.find_spec("nvidia")
    for root in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(root, "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            os.environ["RB_NCCL_LIB"] = cand
            return


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python -c 'import __graft_entry__ as g; g.build()').  reductive_b200 has no CPU fallback.")
    _point_at_bundled_nccl()
    lib = C.CDLL(LIB_PATH)
    sz, pd, vp, fp = C.c_size_t, C.c_ssize_t, C.c_void_p, C.c_void_p
    lib.rb_last_error_message.restype = C.c_char_p
    lib.rb_abi_version.restype = C.c_int
    lib.rb_kernel_launch_count.restype = C.c_uint64
    lib.rb_set_encode_algo.argtypes = [C.c_int]
    lib.rb_set_kmeans_update.argtypes = [C.c_int]
    lib.rb_set_project_algo.argtypes = [C.c_int]
    lib.rb_pq_create.argtypes = [fp, sz, sz, sz, fp, C.POINTER(vp)]
    lib.rb_pq_create_multi.argtypes = [fp, sz, sz, sz, fp, C.POINTER(C.c_int), C.c_int, C.POINTER(vp)]
    lib.rb_pq_n_devices.argtypes = [vp]
    lib.rb_set_host_copy_threads.argtypes = [C.c_int]
    lib.rb_pq_destroy.argtypes = [vp]
    lib.rb_pq_destroy.restype = None
    lib.rb_qstore_create.argtypes = [vp, vp, sz, pd, fp, C.c_int, vp, C.POINTER(vp)]
    lib.rb_qstore_destroy.argtypes = [vp]
    lib.rb_qstore_destroy.restype = None
    lib.rb_qstore_len.argtypes = [vp]
    lib.rb_qstore_len.restype = sz
    lib.rb_qstore_has_norms.argtypes = [vp]
    lib.rb_qstore_embeddings.argtypes = [vp, vp, sz, fp, pd, pd, C.c_int, vp]
    lib.rb_qstore_dot.argtypes = [vp, fp, sz, pd, pd, fp, pd, C.c_int, vp]
    for name in ("rb_pq_quantized_len", "rb_pq_reconstructed_len", "rb_pq_n_quantizer_centroids"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = sz
    lib.rb_pq_has_projection.argtypes = [vp]
    lib.rb_pq_subquantizers.argtypes = [vp, fp]
    lib.rb_pq_projection.argtypes = [vp, fp]
    lib.rb_pq_quantize_batch.argtypes = [vp, fp, sz, pd, pd, vp, C.c_int, pd, pd, C.c_int, vp]
    lib.rb_pq_quantize_vector.argtypes = [vp, fp, pd, vp, C.c_int, pd, C.c_int, vp]
    lib.rb_pq_reconstruct_batch.argtypes = [vp, vp, C.c_int, sz, pd, pd, fp, pd, pd, C.c_int, vp]
    lib.rb_pq_reconstruct.argtypes = [vp, vp, C.c_int, pd, fp, pd, C.c_int, vp]
    lib.rb_check_quantizer_invariants.argtypes = [sz, C.c_uint32, sz, sz, sz, sz, C.POINTER(C.c_uint64)]
    lib.rb_kmeans_packed_len.argtypes = [sz, sz, sz]
    lib.rb_kmeans_packed_len.restype = sz
    lib.rb_kmeans_assign_accumulate.argtypes = [fp, sz, pd, fp, sz, sz, sz, fp, vp]
    lib.rb_kmeans_assign_accumulate_from.argtypes = [fp, sz, pd, fp, sz, sz, sz, fp, fp, vp]
    lib.rb_kmeans_code_pitch.argtypes = [sz]
    lib.rb_kmeans_code_pitch.restype = sz
    lib.rb_kmeans_code_width.argtypes = [sz]
    lib.rb_kmeans_assign.argtypes = [fp, sz, pd, fp, sz, sz, sz, vp, vp]
    lib.rb_kmeans_accumulate.argtypes = [fp, sz, pd, vp, sz, sz, sz, fp, fp, vp]
    lib.rb_kmeans_finalize.argtypes = [fp, sz, sz, sz, C.c_uint64, fp, fp, vp]
    lib.rb_pq_train.argtypes = [fp, sz, sz, pd, pd, sz, C.c_uint32, sz, sz, fp, fp, C.c_int, vp, C.POINTER(vp)]
    lib.rb_project_rows.argtypes = [fp, sz, sz, pd, pd, fp, C.c_int, fp, vp]
    lib.rb_covariance.argtypes = [fp, sz, sz, pd, fp, vp]
    lib.rb_set_gram_algo.argtypes = [C.c_int]
    lib.rb_opq_train_iteration.argtypes = [fp, sz, sz, pd, fp, fp, sz, sz, fp, vp]
    lib.rb_dist_subquantizer_range.argtypes = [sz, C.c_int, C.c_int, C.POINTER(sz), C.POINTER(sz)]
    lib.rb_comm_unique_id.argtypes = [vp, sz]
    lib.rb_comm_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(vp)]
    lib.rb_comm_destroy.argtypes = [vp]
    lib.rb_comm_destroy.restype = None
    lib.rb_comm_rank.argtypes = [vp]
    lib.rb_comm_world.argtypes = [vp]
    lib.rb_kmeans_dist_create.argtypes = [vp, fp, sz, pd, sz, sz, sz, vp, C.POINTER(vp)]
    lib.rb_kmeans_dist_iterate.argtypes = [vp, fp, fp, vp]
    lib.rb_kmeans_dist_destroy.argtypes = [vp]
    lib.rb_kmeans_dist_peer_window.argtypes = [vp]
    lib.rb_kmeans_dist_destroy.restype = None
    lib.rb_pq_train_dist.argtypes = [vp, fp, sz, sz, sz, pd, sz, C.c_uint32, sz, sz, fp, fp, C.c_int, vp, C.POINTER(vp)]
    lib.rb_pq_train_multi.argtypes = [C.POINTER(C.c_int), C.c_int, fp, sz, sz, pd, sz, C.c_uint32, sz, sz, fp, fp,
                                      C.POINTER(vp)]
    return lib


lib = _load()


def last_error() -> str:
    return (lib.rb_last_error_message() or b"").decode("utf-8", "replace")


def check(status: int) -> None:
    """Turn an rb_status into the exception the reference's behaviour maps to."""
    if status == OK:
        return
    msg = last_error()
    if status in _ERR_CLASSES:
        raise _ERR_CLASSES[status](msg)
    if status in (ERR_SHAPE, ERR_CODE_TYPE, ERR_K_MEANS_K):
        raise ReductivePanic(msg)
    if status == ERR_CODE_RANGE:
        raise IndexError(msg)
    if status == ERR_INVALID:
        raise ValueError(msg)
    if status == ERR_NO_DEVICE:
        raise NoDeviceError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if status == ERR_NCCL:
        raise CudaError(f"NCCL: {msg}")
    raise CudaError(f"rb_status {status}: {msg}")


def kernel_launch_count() -> int:
    return int(lib.rb_kernel_launch_count())


def set_encode_algo(algo: int) -> None:
    check(lib.rb_set_encode_algo(algo))


def release_scratch() -> None:
    """Hand the library's cached scratch memory (current device) back to the driver."""
    check(lib.rb_release_scratch())


def set_project_algo(algo: int) -> None:
    """PROJECT_AUTO / PROJECT_EXACT (reference-order FP32 GEMM) / PROJECT_TENSOR (tcgen05, codes still bit-exact)."""
    check(lib.rb_set_project_algo(algo))


GRAM_AUTO, GRAM_CUDA_CORES, GRAM_TENSOR = 0, 1, 2


def set_gram_algo(algo: int) -> None:
    """Gram matrices of Opq training: GRAM_AUTO (tensor cores for n >= 4096, 16 <= d <= 512), GRAM_CUDA_CORES (FP32
    FMA), GRAM_TENSOR (tcgen05, two-limb BF16; UNSUPPORTED outside its range)."""
    check(lib.rb_set_gram_algo(algo))


def set_kmeans_update(ordered) -> None:
    """True / 1 (default): reference summation order, bit-exact sums (stable sort + chain kernels); 3: the same bits
    from the streaming slab kernel in training loops (experimental, slower); False / 0: atomics."""
    check(lib.rb_set_kmeans_update(int(ordered)))
