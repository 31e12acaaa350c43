"""Caller-side quantized storage over the C ABI.

finalfusion stores a trained `Pq<f32>` as a *quantized array*: the quantizer, the `[n, M]` u8 codes from
`QuantizeVector::quantize_batch` (src/pq/traits.rs:77-87) and, optionally, one norm per row (the rows are
l2-normalised before quantisation).  Its lookups are built on `Reconstruct` (src/pq/traits.rs:102-156,
src/pq/pq.rs:303-347):

    embedding(i) = reconstruct(codes[i]) * norms[i]

That type lives in the finalfusion crate, outside /root/reference; `QuantizedArray` keeps its names.  Codes and norms
stay resident in HBM (`rb_qstore_*`, include/reductive_b200.h); `dot` is the fused decode + dot kernel that scores
queries against every stored row without materialising the reconstruction.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from ._cabi import MEM_DEVICE, MEM_HOST, ReductivePanic, check, lib
from .pq import Pq, _Arr, _is_torch


class QuantizedArray:
    def __init__(self, quantizer: Pq, quantized_embeddings, norms=None):
        """quantized_embeddings: [n, M] uint8 (numpy, or a torch CUDA tensor); norms: [n] float32 or None."""
        qa = _Arr(quantized_embeddings, want_float=False)
        if len(qa.shape) != 2 or qa.itemsize != 1:
            raise TypeError("quantized_embeddings must be an [n, M] uint8 matrix")
        if qa.shape[1] != quantizer.quantized_len():
            raise ReductivePanic("Quantization length does not match number of subquantizers")  # primitives.rs:123-127
        if qa.shape[0] > 1 and qa.strides[1] != 1:
            raise ValueError("code rows must be contiguous")
        nptr = None
        if norms is not None:
            na = _Arr(norms, want_float=True)
            if na.shape != (qa.shape[0],):
                raise ReductivePanic("norms must hold one value per row")
            if na.mem != qa.mem:
                raise ValueError("codes and norms must live in the same memory kind")
            if _is_torch(norms):
                norms = norms.contiguous()
            else:
                norms = np.ascontiguousarray(norms, np.float32)
            na = _Arr(norms, want_float=True)
            nptr = na.ptr
        self._quantizer = quantizer  # the store borrows the quantizer handle: keep it alive
        self._h = C.c_void_p()
        row_stride = qa.strides[0] if qa.shape[0] > 1 else qa.shape[1]
        check(lib.rb_qstore_create(quantizer._h, qa.ptr, qa.shape[0], row_stride, nptr, qa.mem, qa.stream, C.byref(self._h)))

    @classmethod
    def quantize_using(cls, quantizer: Pq, embeddings, normalize: bool = True) -> "QuantizedArray":
        """finalfusion's `quantize_using` for a trained quantizer: optionally l2-normalise the rows (keeping the norms),
        then `quantize_batch` (traits.rs:77-87) on the GPU.  The normalisation is caller-side preparation on the host."""
        x = np.ascontiguousarray(embeddings, np.float32)
        norms = None
        if normalize:
            norms = np.sqrt(np.einsum("ij,ij->i", x, x)).astype(np.float32)
            x = x / np.where(norms == 0.0, np.float32(1.0), norms)[:, None]
        codes = quantizer.quantize_batch(x, np.uint8)
        return cls(quantizer, codes, norms)

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib.rb_qstore_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(lib.rb_qstore_len(self._h))

    @property
    def shape(self):
        return (len(self), self._quantizer.reconstructed_len())

    def quantizer(self) -> Pq:
        return self._quantizer

    def has_norms(self) -> bool:
        return bool(lib.rb_qstore_has_norms(self._h))

    def embedding(self, idx: int) -> np.ndarray:
        """`Storage::embedding`: the reconstruction of row idx times its norm."""
        if not 0 <= idx < len(self):
            raise IndexError(idx)
        return self.embeddings(np.array([idx], np.uint64))[0]

    def embeddings(self, indices):
        """`Storage::embeddings`: rows `indices` (numpy -> numpy, torch CUDA int64 tensor -> torch CUDA tensor)."""
        d = self._quantizer.reconstructed_len()
        if _is_torch(indices):
            import torch

            idx = indices.to(torch.int64).contiguous()
            if idx.numel() and (int(idx.min()) < 0 or int(idx.max()) >= len(self)):
                raise IndexError("index out of bounds")
            out = torch.empty((idx.numel(), d), dtype=torch.float32, device=idx.device)
            stream = torch.cuda.current_stream(idx.device).cuda_stream if idx.is_cuda else None
            mem = MEM_DEVICE if idx.is_cuda else MEM_HOST
            check(lib.rb_qstore_embeddings(self._h, idx.data_ptr(), idx.numel(), out.data_ptr(), d, 1, mem, stream))
            return out
        idx = np.ascontiguousarray(indices)
        if idx.size and (idx.min() < 0 or idx.max() >= len(self)):
            raise IndexError("index out of bounds")
        idx = idx.astype(np.uint64)
        out = np.empty((idx.size, d), np.float32)
        check(lib.rb_qstore_embeddings(self._h, idx.ctypes.data, idx.size, out.ctypes.data, d, 1, MEM_HOST, None))
        return out

    def dot(self, queries, out=None):
        """Scores of every stored row against every query: out[q, i] = queries[q] . embedding(i) (fused decode + dot)."""
        qa = _Arr(queries, want_float=True)
        if len(qa.shape) != 2 or qa.shape[1] != self._quantizer.reconstructed_len():
            raise ReductivePanic("Quantizer and vector length mismatch")
        n = len(self)
        if out is None:
            if qa.mem == MEM_DEVICE:
                import torch

                out = torch.empty((qa.shape[0], n), dtype=torch.float32, device=queries.device)
            else:
                out = np.empty((qa.shape[0], n), np.float32)
        oa = _Arr(out, want_float=False)
        if not oa.is_float32 or oa.shape != (qa.shape[0], n) or (n > 1 and oa.strides[1] != 1):
            raise ReductivePanic(f"scores must be a float32 [{qa.shape[0]}, {n}] matrix with contiguous rows")
        if oa.mem != qa.mem:
            raise ValueError("queries and scores must live in the same memory kind")
        row_stride = oa.strides[0] if qa.shape[0] > 1 else n
        check(lib.rb_qstore_dot(self._h, qa.ptr, qa.shape[0], qa.strides[0], qa.strides[1], oa.ptr, row_stride, qa.mem, qa.stream))
        return out


__all__ = ["QuantizedArray"]
