"""Host-side mirror of reductive's quantizer API over the C ABI.

Same names, argument meaning and error behaviour as the reference (paths relative to /root/reference):
  traits  TrainPq          src/pq/traits.rs:15-72
          QuantizeVector   src/pq/traits.rs:75-99
          Reconstruct      src/pq/traits.rs:102-156
  type    Pq               src/pq/pq.rs:29-32 (+ impls :196-348)
`ndarray` views become numpy arrays (host memory, any strides) or torch CUDA tensors (device memory of the
current device, any strides); the index type `I` becomes a numpy integer dtype.  Where the reference panics
(assert!), these raise ReductivePanic / IndexError; where it returns Err(ReductiveError::X), they raise X.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _cabi
from ._cabi import MEM_DEVICE, MEM_HOST, ReductivePanic, check, lib

_CODE_DTYPES = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}


def _is_torch(a) -> bool:
    return type(a).__module__.split(".")[0] == "torch"


class _Arr:
    """Pointer + shape + element strides + memory kind of a numpy array or torch tensor."""

    __slots__ = ("ptr", "shape", "strides", "mem", "stream", "keep", "itemsize", "is_float32")

    def __init__(self, a, want_float: bool):
        if _is_torch(a):
            import torch

            if want_float and a.dtype != torch.float32:
                raise TypeError("expected a float32 tensor")
            self.ptr = a.data_ptr()
            self.shape = tuple(a.shape)
            self.strides = tuple(a.stride())
            self.itemsize = a.element_size()
            self.is_float32 = a.dtype == torch.float32
            if a.is_cuda:
                self.mem = MEM_DEVICE
                self.stream = torch.cuda.current_stream(a.device).cuda_stream
            else:
                self.mem = MEM_HOST
                self.stream = None
            self.keep = a
        else:
            a = np.asarray(a)
            if want_float and a.dtype != np.float32:
                a = a.astype(np.float32)
            self.ptr = a.ctypes.data
            self.shape = a.shape
            self.itemsize = a.itemsize
            self.strides = tuple(s // a.itemsize for s in a.strides)
            self.is_float32 = a.dtype == np.float32
            self.mem = MEM_HOST
            self.stream = None
            self.keep = a


def _code_width(dtype) -> int:
    dt = np.dtype(dtype)
    if dt.kind not in "ui" or dt.itemsize not in _CODE_DTYPES:
        raise TypeError(f"unsupported code dtype {dt}")
    return dt.itemsize


def _torch_code_dtype(width: int):
    import torch

    return {1: torch.uint8, 2: torch.int16, 4: torch.int32, 8: torch.int64}[width]


class QuantizeVector:
    """Vector quantization (trait QuantizeVector, src/pq/traits.rs:75-99)."""

    def quantize_batch(self, x, dtype=np.uint8):
        raise NotImplementedError

    def quantize_batch_into(self, x, quantized) -> None:
        raise NotImplementedError

    def quantize_vector(self, x, dtype=np.uint8):
        raise NotImplementedError

    def quantized_len(self) -> int:
        raise NotImplementedError


class Reconstruct:
    """Vector reconstruction (trait Reconstruct, src/pq/traits.rs:102-156)."""

    def reconstruct_batch(self, quantized):  # provided method, traits.rs:109-117
        q = _Arr(quantized, want_float=False)
        if len(q.shape) != 2:
            raise ReductivePanic("quantized must be a matrix")
        n = q.shape[0]
        if q.mem == MEM_DEVICE:
            import torch

            out = torch.zeros((n, self.reconstructed_len()), dtype=torch.float32, device=quantized.device)
        else:
            out = np.zeros((n, self.reconstructed_len()), np.float32)
        self.reconstruct_batch_into(quantized, out)
        return out

    def reconstruct_batch_into(self, quantized, reconstructions) -> None:
        raise NotImplementedError

    def reconstruct(self, quantized):  # provided method, traits.rs:133-141
        q = _Arr(quantized, want_float=False)
        if q.mem == MEM_DEVICE:
            import torch

            out = torch.zeros((self.reconstructed_len(),), dtype=torch.float32, device=quantized.device)
        else:
            out = np.zeros((self.reconstructed_len(),), np.float32)
        self.reconstruct_into(quantized, out)
        return out

    def reconstruct_into(self, quantized, reconstruction) -> None:
        raise NotImplementedError

    def reconstructed_len(self) -> int:
        raise NotImplementedError


class TrainPq:
    """Training trait for product quantizers (trait TrainPq, src/pq/traits.rs:15-72)."""

    @classmethod
    def train_pq(cls, n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, instances):
        # traits.rs:26-44: seeds a ChaCha8Rng from entropy
        return cls.train_pq_using(n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, instances,
                                  np.random.default_rng())

    @classmethod
    def train_pq_using(cls, n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, instances, rng):
        raise NotImplementedError


def random_instance_centroids(instances_shape: Tuple[int, int], n_subquantizers: int, k: int, n_attempts: int,
                              rng: np.random.Generator) -> np.ndarray:
    """Row indices of the initial centroids, [n_attempts, M, k]: k distinct uniform rows per subquantizer and
    attempt (RandomInstanceCentroids, src/kmeans.rs:52-87; one derived RNG per subquantizer like the
    XorShiftRng split at src/pq/pq.rs:221-224)."""
    n = instances_shape[0]
    if k == 0:
        raise ReductivePanic("Cannot pick 0 random centroids")  # kmeans.rs:61
    if k >= n:
        raise ReductivePanic(  # kmeans.rs:62-67
            f"Cannot pick more centroids than instances: {n} instances, {k} centroids")
    seeds = rng.integers(0, 2 ** 63 - 1, size=n_subquantizers)
    idx = np.empty((n_attempts, n_subquantizers, k), np.int64)
    for m in range(n_subquantizers):
        sub = np.random.default_rng(int(seeds[m]))
        for a in range(n_attempts):
            idx[a, m] = sub.choice(n, size=k, replace=False)
    return idx


def check_quantizer_invariants(n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, n_rows, n_cols):
    """Pq::check_quantizer_invariants (src/pq/pq.rs:63-100); raises the matching ReductiveError."""
    detail = C.c_uint64(0)
    check(lib.rb_check_quantizer_invariants(n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts,
                                            n_rows, n_cols, C.byref(detail)))


class Pq(QuantizeVector, Reconstruct, TrainPq):
    """Product quantizer (Jégou et al., 2011) resident on the current CUDA device.

    Mirror of reductive::pq::Pq<f32> (src/pq/pq.rs:29-32)."""

    def __init__(self, projection: Optional[np.ndarray], quantizers: np.ndarray,
                 devices: Optional[Sequence[int]] = None):
        # Pq::new, pq.rs:38-61.  `devices`: replicate the quantizer on several GPUs of this process
        # (rb_pq_create_multi); host-memory batches are then split over them.
        q = np.ascontiguousarray(quantizers, np.float32)
        if q.ndim != 3 or q.size == 0:
            raise ReductivePanic("Attempted to construct a product quantizer without quantizers.")
        M, k, dsub = q.shape
        p = None
        if projection is not None:
            p = np.ascontiguousarray(projection, np.float32)
            if p.shape != (M * dsub, M * dsub):
                raise ReductivePanic(
                    f"Incorrect projection matrix shape, was: {list(p.shape)}, should be [{M * dsub}, {M * dsub}]")
        self._h = C.c_void_p()
        if devices is not None and len(devices) > 0:
            dev = (C.c_int * len(devices))(*[int(v) for v in devices])
            check(lib.rb_pq_create_multi(q.ctypes.data, M, k, dsub, None if p is None else p.ctypes.data, dev,
                                         len(devices), C.byref(self._h)))
        else:
            check(lib.rb_pq_create(q.ctypes.data, M, k, dsub, None if p is None else p.ctypes.data, C.byref(self._h)))

    @classmethod
    def _from_handle(cls, handle: C.c_void_p) -> "Pq":
        self = cls.__new__(cls)
        self._h = handle
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            lib.rb_pq_destroy(h)
            self._h = None

    # ---- accessors ------------------------------------------------------------------------------
    def n_quantizer_centroids(self) -> int:  # pq.rs:103
        return int(lib.rb_pq_n_quantizer_centroids(self._h))

    def quantized_len(self) -> int:  # pq.rs:300
        return int(lib.rb_pq_quantized_len(self._h))

    def reconstructed_len(self) -> int:  # pq.rs:345
        return int(lib.rb_pq_reconstructed_len(self._h))

    def subquantizers(self) -> np.ndarray:  # pq.rs:191
        M, k = self.quantized_len(), self.n_quantizer_centroids()
        out = np.empty((M, k, self.reconstructed_len() // M), np.float32)
        check(lib.rb_pq_subquantizers(self._h, out.ctypes.data))
        return out

    def projection(self) -> Optional[np.ndarray]:  # pq.rs:108
        if not lib.rb_pq_has_projection(self._h):
            return None
        d = self.reconstructed_len()
        out = np.empty((d, d), np.float32)
        check(lib.rb_pq_projection(self._h, out.ctypes.data))
        return out

    def __eq__(self, other):  # #[derive(PartialEq)], pq.rs:28
        if not isinstance(other, Pq):
            return NotImplemented
        pa, pb = self.projection(), other.projection()
        return (np.array_equal(self.subquantizers(), other.subquantizers())
                and ((pa is None and pb is None) or (pa is not None and pb is not None and np.array_equal(pa, pb))))

    # ---- TrainPq (pq.rs:196-250) ------------------------------------------------------------------
    @classmethod
    def train_pq_using(cls, n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, instances, rng,
                       initial_centroids: Optional[np.ndarray] = None, return_loss: bool = False,
                       devices: Optional[Sequence[int]] = None):
        """`initial_centroids` ([n_attempts, M, k, dsub]) overrides the random instance draw — the hook the
        parity tests use to start the CUDA path and the oracle from identical centroids.  `devices` (host
        instances only): train on several GPUs from this one process (rb_pq_train_multi: rows split into contiguous
        blocks, bit-identical to the one-GPU result); the quantizer lives on devices[0]."""
        x = _Arr(instances, want_float=True)
        if len(x.shape) != 2:
            raise ReductivePanic("instances must be a matrix")
        n, d = x.shape
        check_quantizer_invariants(n_subquantizers, n_subquantizer_bits, n_iterations, n_attempts, n, d)
        M, k = n_subquantizers, 1 << n_subquantizer_bits
        dsub = d // M
        if initial_centroids is None:
            idx = random_instance_centroids((n, d), M, k, n_attempts, rng)
            rows = np.unique(idx)
            if x.mem == MEM_DEVICE:
                import torch

                picked = instances[torch.as_tensor(rows, device=instances.device)].cpu().numpy()
            else:
                picked = np.asarray(x.keep)[rows]
            pos = np.searchsorted(rows, idx)  # [A, M, k] -> position in `picked`
            init = np.empty((n_attempts, M, k, dsub), np.float32)
            for m in range(M):
                init[:, m] = picked[pos[:, m]][:, :, m * dsub:(m + 1) * dsub]
        else:
            init = np.ascontiguousarray(initial_centroids, np.float32).reshape(n_attempts, M, k, dsub)
        loss = np.zeros((M,), np.float32)
        h = C.c_void_p()
        if devices is not None and len(devices) > 1:
            if x.mem != MEM_HOST or x.strides[1] != 1:
                raise ValueError("multi-device training takes host instances with unit column stride")
            dev = (C.c_int * len(devices))(*[int(v) for v in devices])
            check(lib.rb_pq_train_multi(dev, len(devices), x.ptr, n, d, x.strides[0], M, n_subquantizer_bits,
                                        n_iterations, n_attempts, init.ctypes.data, loss.ctypes.data, C.byref(h)))
            pq = cls._from_handle(h)
            return (pq, loss) if return_loss else pq
        check(lib.rb_pq_train(x.ptr, n, d, x.strides[0], x.strides[1], M, n_subquantizer_bits, n_iterations,
                              n_attempts, init.ctypes.data, loss.ctypes.data, x.mem, x.stream, C.byref(h)))
        pq = cls._from_handle(h)
        return (pq, loss) if return_loss else pq

    # ---- QuantizeVector (pq.rs:252-303) -----------------------------------------------------------
    def quantize_batch(self, x, dtype=np.uint8):
        xa = _Arr(x, want_float=True)
        if len(xa.shape) != 2:
            raise ReductivePanic("x must be a matrix")
        width = _code_width(dtype)
        if xa.mem == MEM_DEVICE:
            import torch

            out = torch.zeros((xa.shape[0], self.quantized_len()), dtype=_torch_code_dtype(width), device=x.device)
        else:
            out = np.zeros((xa.shape[0], self.quantized_len()), dtype)
        self.quantize_batch_into(x, out)
        return out

    def quantize_batch_into(self, x, quantized) -> None:
        xa = _Arr(x, want_float=True)
        qa = _Arr(quantized, want_float=False)
        if len(xa.shape) != 2 or len(qa.shape) != 2:
            raise ReductivePanic("x and quantized must be matrices")
        if xa.shape[1] != self.reconstructed_len():  # primitives.rs:74-78
            raise ReductivePanic("Quantizer and vector length mismatch")
        if qa.shape != (xa.shape[0], self.quantized_len()):  # primitives.rs:80-87
            raise ReductivePanic(
                f"Quantized matrix has incorrect shape, expected: ({xa.shape[0]}, {self.quantized_len()}), "
                f"got: ({qa.shape[0]}, {qa.shape[1]})")
        if xa.mem != qa.mem:
            raise ValueError("x and quantized must live in the same memory kind")
        check(lib.rb_pq_quantize_batch(self._h, xa.ptr, xa.shape[0], xa.strides[0], xa.strides[1], qa.ptr,
                                       qa.itemsize, qa.strides[0], qa.strides[1], xa.mem, xa.stream))

    def quantize_vector(self, x, dtype=np.uint8):
        xa = _Arr(x, want_float=True)
        if len(xa.shape) != 1 or xa.shape[0] != self.reconstructed_len():  # primitives.rs:25-29
            raise ReductivePanic("Quantizer and vector length mismatch")
        width = _code_width(dtype)
        if xa.mem == MEM_DEVICE:
            import torch

            out = torch.zeros((self.quantized_len(),), dtype=_torch_code_dtype(width), device=x.device)
            optr, ostride = out.data_ptr(), 1
        else:
            out = np.zeros((self.quantized_len(),), dtype)
            optr, ostride = out.ctypes.data, 1
        check(lib.rb_pq_quantize_vector(self._h, xa.ptr, xa.strides[0], optr, width, ostride, xa.mem, xa.stream))
        return out

    # ---- Reconstruct (pq.rs:305-348) --------------------------------------------------------------
    def reconstruct_batch_into(self, quantized, reconstructions) -> None:
        qa = _Arr(quantized, want_float=False)
        ra = _Arr(reconstructions, want_float=False)
        if not ra.is_float32:
            raise TypeError("reconstructions must be float32")
        if len(qa.shape) != 2 or len(ra.shape) != 2:
            raise ReductivePanic("quantized and reconstructions must be matrices")
        if qa.shape[1] != self.quantized_len():  # primitives.rs:123-127
            raise ReductivePanic("Quantization length does not match number of subquantizers")
        if ra.shape != (qa.shape[0], self.reconstructed_len()):  # primitives.rs:159-167
            raise ReductivePanic(
                f"Reconstructions matrix has incorrect shape, expected: ({qa.shape[0]}, {self.reconstructed_len()}), "
                f"got: ({ra.shape[0]}, {ra.shape[1]})")
        if qa.mem != ra.mem:
            raise ValueError("quantized and reconstructions must live in the same memory kind")
        check(lib.rb_pq_reconstruct_batch(self._h, qa.ptr, qa.itemsize, qa.shape[0], qa.strides[0], qa.strides[1],
                                          ra.ptr, ra.strides[0], ra.strides[1], qa.mem, qa.stream))

    def reconstruct_into(self, quantized, reconstruction) -> None:
        qa = _Arr(quantized, want_float=False)
        ra = _Arr(reconstruction, want_float=False)
        if not ra.is_float32:
            raise TypeError("reconstruction must be float32")
        if len(qa.shape) != 1 or qa.shape[0] != self.quantized_len():  # primitives.rs:123-127
            raise ReductivePanic("Quantization length does not match number of subquantizers")
        if len(ra.shape) != 1 or ra.shape[0] != self.reconstructed_len():  # primitives.rs:129-135
            raise ReductivePanic(
                f"Reconstructed output length ({ra.shape[0] if ra.shape else 0}) does not match reconstructed "
                f"vector length ({self.reconstructed_len()})")
        if qa.mem != ra.mem:
            raise ValueError("quantized and reconstruction must live in the same memory kind")
        check(lib.rb_pq_reconstruct(self._h, qa.ptr, qa.itemsize, qa.strides[0], ra.ptr, ra.strides[0], qa.mem,
                                    qa.stream))
