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
    """norm[i] = SUM_f SUM_k a_f[k, i] * b_f[k, i]  (fp64), one kernel for up to 32 field pairs."""
    lib = _lib.load()
    dims = _dims_of(a[0])
    dev = a[0].buffer.device
    n = len(a)
    if n > 32 or n != len(b):
        raise ValueError("symmetry_norms: 1..32 field pairs")
    total = torch.empty(dims.ncol, dtype=torch.float64, device=dev) if out is None else out
    _lib.check(
        lib.cs2_symmetry_norms(C.byref(dims), n, _ptr_array(a, n), _ptr_array(b, n), total.data_ptr(),
                               torch.cuda.current_stream(dev).cuda_stream),
        "cs2_symmetry_norms",
    )
    return total


class SymmetryResidual:
    """norm3[i] = |n1 - n2| / eps where n2 == 0 else |n1 - n2| / (eps n2), and max_i norm3[i], on the device
    (adjoint/validation.py:157-165).  `max` is a 1-element fp64 device tensor: what a sharded run all-reduces."""

    def __init__(self) -> None:
        self.lib = _lib.load()
        self._scratch: Optional[torch.Tensor] = None
        self.norm3: Optional[torch.Tensor] = None
        self.max: Optional[torch.Tensor] = None

    def __call__(self, norm1: torch.Tensor, norm2: torch.Tensor, eps: float):
        dev, n = norm1.device, norm1.numel()
        if dev.type != "cuda":
            raise _lib.CUDAExtensionError("reductions run only as CUDA kernels")
        assert norm1.dtype == norm2.dtype == torch.float64 and norm2.numel() == n and norm2.device == dev
        nbytes = self.lib.cs2_symmetry_residual_scratch_bytes(n)
        if self._scratch is None or self._scratch.numel() < nbytes or self._scratch.device != dev:
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.max = torch.empty(1, dtype=torch.float64, device=dev)
        if self.norm3 is None or self.norm3.numel() != n or self.norm3.device != dev:
            self.norm3 = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.check(
            self.lib.cs2_symmetry_residual(n, norm1.data_ptr(), norm2.data_ptr(), float(eps), self.norm3.data_ptr(),
                                           self.max.data_ptr(), self._scratch.data_ptr(), self._scratch.numel(),
                                           torch.cuda.current_stream(dev).cuda_stream),
            "cs2_symmetry_residual",
        )
        return self.norm3, self.max
