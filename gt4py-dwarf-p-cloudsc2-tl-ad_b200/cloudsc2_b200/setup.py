"""State construction (reference: cloudsc2_gt4py/setup.py:28-70).

`get_state(hdf5_grid_operator)` loads the 16 named fields from an `input.h5`-style file (2-D
datasets `(K, IJ)`, 5-species datasets `(5, K, IJ)`), tiling the file's KLON columns to the
grid's nx (column i <- column i mod KLON).  `get_synthetic_state` produces the same dictionary
from the seeded generator in `synthetic.py` (the reference's `input.h5` is not shipped).
"""
from __future__ import annotations

from datetime import datetime
from typing import Any, Dict, Optional

import numpy as np

from . import synthetic
from .framework.config import GT4PyConfig
from .framework.grid import ComputationalGrid, I, J, K
from .framework.storage import Field, zeros
from .h5lite import File

REFERENCE_TIME = datetime(year=1970, month=1, day=1)

# state key -> (HDF5 dataset, species index or None, half-level?, units)   (setup.py:48-65)
FIELD_PROPERTIES = {
    "f_a": ("PA", None, False, "1"),
    "f_ap": ("PAP", None, False, "Pa"),
    "f_aph": ("PAPH", None, True, "Pa"),
    "f_lu": ("PLU", None, False, "g g^-1"),
    "f_lude": ("PLUDE", None, False, "kg m^-3 s^-1"),
    "f_mfd": ("PMFD", None, False, "kg m^-2 s^-1"),
    "f_mfu": ("PMFU", None, False, "kg m^-2 s^-1"),
    "f_qi": ("PCLV", 1, False, "g g^-1"),
    "f_ql": ("PCLV", 0, False, "g g^-1"),
    "f_q": ("PQ", None, False, "g g^-1"),
    "f_supsat": ("PSUPSAT", None, False, "g g^-1"),
    "f_t": ("PT", None, False, "K"),
    "f_tnd_cml_qi": ("TENDENCY_CML_CLD", 1, False, "g g^-1 s^-1"),
    "f_tnd_cml_ql": ("TENDENCY_CML_CLD", 0, False, "g g^-1 s^-1"),
    "f_tnd_cml_q": ("TENDENCY_CML_Q", None, False, "g g^-1 s^-1"),
    "f_tnd_cml_t": ("TENDENCY_CML_T", None, False, "K s^-1"),
}


class HDF5GridOperator:
    """Stand-in for `ifs_physics_common.iox.HDF5GridOperator`: reads a dataset and returns it as a
    `Field` on the computational grid, replicating columns when nx > KLON."""

    def __init__(self, filename: str, computational_grid: ComputationalGrid, *, gt4py_config: GT4PyConfig,
                 column_offset: int = 0) -> None:
        self.f = File(filename)
        self.computational_grid = computational_grid
        self.gt4py_config = gt4py_config
        self.column_offset = column_offset  # first GLOBAL column of this rank's shard

    def get_field(self, h5_name: str, index: Optional[int] = None, half: bool = False, units: str = "", name: str = "") -> Field:
        data = self.f[h5_name]
        if index is not None:
            data = data[index]
        nx = self.computational_grid.nx
        cols = (np.arange(nx) + self.column_offset) % data.shape[1]
        dims = (I, J, K - 1 / 2) if half else (I, J, K)
        fld = zeros(self.computational_grid, dims, gt4py_config=self.gt4py_config, units=units, name=name)
        return fld.assign(data[:, cols])


def get_state(hdf5_grid_operator: HDF5GridOperator) -> Dict[str, Any]:
    state: Dict[str, Any] = {}
    for name, (h5_name, index, half, units) in FIELD_PROPERTIES.items():
        # a missing dataset raises KeyError, like the reference's get_field (setup.py:66-68)
        state[name] = hdf5_grid_operator.get_field(h5_name, index, half, units, name)
    state["time"] = REFERENCE_TIME
    return state


def state_from_arrays(arrays: Dict[str, np.ndarray], computational_grid: ComputationalGrid, *, gt4py_config: GT4PyConfig) -> Dict[str, Any]:
    """Build a state dict from host arrays in `(K, IJ)` orientation (e.g. synthetic.base_block())."""
    state: Dict[str, Any] = {}
    for name, arr in arrays.items():
        half = name == "f_aph"
        units = FIELD_PROPERTIES[name][3] if name in FIELD_PROPERTIES else ""
        dims = (I, J, K - 1 / 2) if half else (I, J, K)
        state[name] = zeros(computational_grid, dims, gt4py_config=gt4py_config, units=units, name=name).assign(arr)
    state["time"] = REFERENCE_TIME
    return state


def get_synthetic_state(computational_grid: ComputationalGrid, *, gt4py_config: GT4PyConfig, block: str = "base",
                        seed: int = 0, column_offset: int = 0) -> Dict[str, Any]:
    """Synthetic state tiled from the 100-column block (column i <- i mod 100, like `--num-cols` tiles input.h5);
    `column_offset` = first global column of a shard.  The block is uploaded once and tiled ON THE DEVICE, so a
    1 M-column state costs no 20 GB of host arrays."""
    import torch

    from .framework.storage import default_device, torch_dtype

    nz, nx = computational_grid.nz, computational_grid.nx
    blk = synthetic.base_block(nz=nz, seed=seed) if block == "base" else synthetic.cold_block(nz=nz, seed=seed + 1)
    dev = default_device(gt4py_config)
    if dev.type != "cuda" or nx <= 4 * synthetic.KLON:
        cols = (np.arange(nx) + column_offset) % synthetic.KLON
        return state_from_arrays({k: v[:, cols] for k, v in blk.items()}, computational_grid, gt4py_config=gt4py_config)
    cols = (torch.arange(nx, device=dev) + column_offset) % synthetic.KLON
    dt = torch_dtype(gt4py_config.dtypes.float)
    state: Dict[str, Any] = {}
    for name, arr in blk.items():
        dims = (I, J, K - 1 / 2) if name == "f_aph" else (I, J, K)
        units = FIELD_PROPERTIES[name][3] if name in FIELD_PROPERTIES else ""
        fld = zeros(computational_grid, dims, gt4py_config=gt4py_config, units=units, name=name)
        src = torch.as_tensor(np.ascontiguousarray(arr), device=dev).to(dt)  # host rounding to the field dtype
        fld.buffer[: src.shape[0], :nx].copy_(src.index_select(1, cols))
        state[name] = fld
    state["time"] = REFERENCE_TIME
    return state
