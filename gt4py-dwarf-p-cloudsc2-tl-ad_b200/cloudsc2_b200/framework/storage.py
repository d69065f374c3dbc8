"""Field storage (stand-in for `ifs_physics_common.storage` + the sympl DataArray).

Layout (DESIGN.md "Data layout in HBM"): every (I, J, K) or (I, J, K-1/2) field is ONE
allocation `buffer[nz+1][ncol_stride]`, column index fastest, `ncol_stride` = nx rounded up to
a multiple of 32 elements; torch's caching allocator returns 512-byte aligned blocks, so every
level row starts 256-byte (fp64) / 128-byte (fp32) aligned.  The logical `(nx, 1, nz+1)` array the
reference harnesses index (`.data[:, 0, :]`, tangent_linear/validation.py:243) is the zero-copy
strided view `buffer[:, :nx].T[:, None, :]`.  K-only fields (f_eta) live on the host.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Any, Dict, Iterator, Optional, Sequence, Tuple

import numpy as np
import torch

from .config import GT4PyConfig
from .grid import ComputationalGrid, DimSymbol, I, J, K

_TORCH_DTYPES = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                 np.dtype(np.int64): torch.int64, np.dtype(np.int32): torch.int32, np.dtype(bool): torch.bool}


def torch_dtype(np_dtype: Any) -> torch.dtype:
    return _TORCH_DTYPES[np.dtype(np_dtype)]


def default_device(gt4py_config: Optional[GT4PyConfig] = None) -> torch.device:
    if gt4py_config is not None and gt4py_config.device is not None:
        return torch.device(gt4py_config.device)
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def column_stride(nx: int) -> int:
    return max(32, -(-nx // 32) * 32)


class Field:
    """DataArray-like wrapper: `.data` (logical view), `.dims`, `.attrs["units"]`."""

    def __init__(self, buffer: torch.Tensor, nx: Optional[int], grid_dims: Tuple[DimSymbol, ...], units: str = "",
                 name: str = "") -> None:
        self.buffer = buffer
        self.nx = nx
        self.grid_dims = tuple(grid_dims)
        self.attrs: Dict[str, Any] = {"units": units}
        self.name = name
        self._view: Optional[torch.Tensor] = None
        self._view_of: Tuple[Any, Any] = (None, None)

    @property
    def data(self) -> torch.Tensor:
        """Logical `(nx, 1, nz+1)` view of the column-fastest buffer.  The view object is built once per (buffer, nx) and
        handed out again: a component call touches ~30 fields, and building the views anew was a third of its host time."""
        buf = self.buffer
        if self._view is None or self._view_of[0] is not buf or self._view_of[1] != self.nx:
            self._view = buf if buf.dim() == 1 else buf[:, : self.nx].t().unsqueeze(1)
            self._view_of = (buf, self.nx)
        return self._view

    @property
    def dims(self) -> Tuple[str, ...]:
        return tuple(repr(d) for d in self.grid_dims)

    @property
    def shape(self) -> Tuple[int, ...]:
        return tuple(self.data.shape)

    @property
    def dtype(self) -> torch.dtype:
        return self.buffer.dtype

    def numpy(self) -> np.ndarray:
        """Host copy in the HDF5 `(K, IJ)` orientation `[nz+1, nx]` (or `[nz+1]` for K-fields)."""
        if self.buffer.dim() == 1:
            return self.buffer.detach().cpu().numpy().copy()
        return self.buffer[:, : self.nx].detach().cpu().numpy().copy()

    def assign(self, array_k_ij: Any) -> "Field":
        """Fill from a host array in `(K, IJ)` orientation; full-level data may omit the padding level."""
        src = torch.as_tensor(np.ascontiguousarray(array_k_ij), dtype=self.buffer.dtype)
        if self.buffer.dim() == 1:
            self.buffer[: src.shape[0]].copy_(src)
        else:
            self.buffer[: src.shape[0], : self.nx].copy_(src, non_blocking=False)
        return self

    def __repr__(self) -> str:
        return f"Field({self.name!r}, dims={self.dims}, shape={self.shape}, dtype={self.dtype}, device={self.buffer.device})"


def zeros(computational_grid: ComputationalGrid, grid_dims: Sequence[DimSymbol], *, gt4py_config: GT4PyConfig,
          dtype_name: str = "float", units: str = "", name: str = "") -> Field:
    """Zero-initialised storage for the given grid dims (like `ifs_physics_common.storage.zeros`)."""
    np_dtype = getattr(gt4py_config.dtypes, dtype_name)
    dt = torch_dtype(np_dtype)
    grid_dims = tuple(grid_dims)
    nx, nz = computational_grid.nx, computational_grid.nz
    if len(grid_dims) == 3:  # (I, J, K) and (I, J, K-1/2) share the nz+1 storage
        buf = torch.zeros((nz + 1, column_stride(nx)), dtype=dt, device=default_device(gt4py_config))
        return Field(buf, nx, grid_dims, units, name)
    if len(grid_dims) == 2:  # (I, J) scratch
        buf = torch.zeros((1, column_stride(nx)), dtype=dt, device=default_device(gt4py_config))
        return Field(buf, nx, grid_dims, units, name)
    if len(grid_dims) == 1:  # (K,) fields are host-resident
        return Field(torch.zeros(nz + 1, dtype=dt, device="cpu"), None, grid_dims, units, name)
    raise ValueError(f"unsupported grid dims {grid_dims}")


def gt_zeros(computational_grid: ComputationalGrid, grid_dims: Sequence[DimSymbol], *, gt4py_config: GT4PyConfig,
             dtype_name: str = "float") -> torch.Tensor:
    """Raw-array flavour used for `klevel` (tangent_linear/microphysics.py:68-71)."""
    return zeros(computational_grid, grid_dims, gt4py_config=gt4py_config, dtype_name=dtype_name).data


@contextmanager
def managed_temporary_storage(computational_grid: ComputationalGrid, *specs: Tuple[Tuple[DimSymbol, ...], str],
                              gt4py_config: GT4PyConfig) -> Iterator[Tuple[Any, ...]]:
    """The reference allocates IJ scratch fields for the flux / overlap carries
    (nonlinear/microphysics.py:131-133).  The B200 kernels keep those carries in registers, so
    nothing is allocated: the context yields `None` placeholders which the stencil callables
    accept and ignore."""
    yield tuple(None for _ in specs)


def allocate_field(computational_grid: ComputationalGrid, name: str, props: Dict[str, Any], gt4py_config: GT4PyConfig) -> Field:
    return zeros(computational_grid, props["grid_dims"], gt4py_config=gt4py_config,
                 dtype_name=props.get("dtype_name", "float"), units=props.get("units", ""), name=name)


def field_from_numpy(computational_grid: ComputationalGrid, array_k_ij: np.ndarray, grid_dims=(I, J, K), *,
                     gt4py_config: GT4PyConfig, units: str = "", name: str = "") -> Field:
    return zeros(computational_grid, grid_dims, gt4py_config=gt4py_config, units=units, name=name).assign(array_k_ij)
