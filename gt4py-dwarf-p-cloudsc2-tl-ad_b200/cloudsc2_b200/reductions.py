"""Device-side reductions of the validation harnesses (C ABI: cs2_taylor_sums,
cs2_symmetry_norms).  They replace the host NumPy sums of the reference
(tangent_linear/validation.py:252-261, adjoint/validation.py:167-215): nothing but a few doubles
ever leaves the GPU, and under column sharding those doubles are all-reduced once per test."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .framework.storage import Field


def _dims_of(fld: Field) -> _lib.Dims:
    buf = fld.buffer
    if buf.device.type != "cuda":
        raise _lib.CUDAExtensionError("reductions run only as CUDA kernels; field lives on " + str(buf.device))
    code = _lib.CS2_F64 if buf.dtype == torch.float64 else _lib.CS2_F32
    return _lib.Dims(fld.nx, buf.shape[1], buf.shape[0] - 1, code)


def _ptr_array(fields: Optional[Sequence[Optional[Field]]], n: int):
    arr = (C.c_void_p * n)()
    for i in range(n):
        f = fields[i] if fields is not None else None
        arr[i] = f.buffer.data_ptr() if f is not None else None
    return arr


class TaylorSums:
    """Accumulates, per field f, SUM(a_f - b_f) and SUM(c_f) over all levels and columns."""

    def __init__(self) -> None:
        self.lib = _lib.load()
        self._scratch: Optional[torch.Tensor] = None

    def __call__(self, a: Sequence[Field], b: Optional[Sequence[Optional[Field]]], c: Optional[Sequence[Optional[Field]]],
                 sums: torch.Tensor) -> torch.Tensor:
        n = len(a)
        dims = _dims_of(a[0])
        nbytes = self.lib.cs2_taylor_scratch_bytes(C.byref(dims), n)
        dev = a[0].buffer.device
        if self._scratch is None or self._scratch.numel() < nbytes or self._scratch.device != dev:
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        assert sums.dtype == torch.float64 and sums.numel() >= 2 * n and sums.device == dev
        _lib.check(
            self.lib.cs2_taylor_sums(C.byref(dims), n, _ptr_array(a, n), _ptr_array(b, n), _ptr_array(c, n),
                                     sums.data_ptr(), self._scratch.data_ptr(), self._scratch.numel(),
                                     torch.cuda.current_stream(dev).cuda_stream),
            "cs2_taylor_sums",
        )
        return sums


def symmetry_norms(a: Sequence[Field], b: Sequence[Field], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """norm[i] = SUM_f SUM_k a_f[k, i] * b_f[k, i]  (fp64), in chunks of <= 16 field pairs."""
    lib = _lib.load()
    dims = _dims_of(a[0])
    dev = a[0].buffer.device
    total = torch.zeros(dims.ncol, dtype=torch.float64, device=dev) if out is None else out.zero_()
    part = torch.empty_like(total)
    for lo in range(0, len(a), 16):
        aa, bb = a[lo : lo + 16], b[lo : lo + 16]
        _lib.check(
            lib.cs2_symmetry_norms(C.byref(dims), len(aa), _ptr_array(aa, len(aa)), _ptr_array(bb, len(bb)),
                                   part.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
            "cs2_symmetry_norms",
        )
        total += part
    return total
