"""Grid description (stand-in for `ifs_physics_common.grid`): dimension symbols `I, J, K`,
the staggered `K - 1/2`, and `ComputationalGrid(GridConfig(nx, ny, nz))` whose
`.grids[dims].shape` the reference components use for the stencil domains
(e.g. nonlinear/microphysics.py:169)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

from .config import GridConfig


@dataclass(frozen=True)
class DimSymbol:
    name: str
    offset: float = 0.0

    def __sub__(self, other: float) -> "DimSymbol":
        return DimSymbol(self.name, self.offset - other)

    def __add__(self, other: float) -> "DimSymbol":
        return DimSymbol(self.name, self.offset + other)

    def __getitem__(self, index: int) -> "DimSymbol":  # D5[index]
        return DimSymbol(f"{self.name}[{index}]", self.offset)

    def __repr__(self) -> str:
        return self.name if not self.offset else f"{self.name}{self.offset:+g}"


I = DimSymbol("I")  # noqa: E741
J = DimSymbol("J")
K = DimSymbol("K")
IJ = DimSymbol("IJ")
D5 = DimSymbol("D5")
ExpandedDim = DimSymbol("ExpandedDim")


@dataclass(frozen=True)
class Grid:
    shape: Tuple[int, ...]
    dims: Tuple[str, ...]


class ComputationalGrid:
    def __init__(self, grid_config: GridConfig) -> None:
        self.grid_config = grid_config
        nx, ny, nz = grid_config.nx, grid_config.ny, grid_config.nz
        if ny != 1:
            raise ValueError("CLOUDSC2 columns are laid out with ny = 1 (drivers/run_nonlinear.py:57)")
        self.nx, self.ny, self.nz = nx, ny, nz
        self.grids: Dict[Tuple[DimSymbol, ...], Grid] = {
            (I, J, K): Grid((nx, ny, nz), ("x", "y", "z")),
            (I, J, K - 1 / 2): Grid((nx, ny, nz + 1), ("x", "y", "z_h")),
            (I, J): Grid((nx, ny), ("x", "y")),
            (K,): Grid((nz,), ("z",)),
            (K - 1 / 2,): Grid((nz + 1,), ("z_h",)),
        }
