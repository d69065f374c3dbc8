"""Configuration objects (stand-ins for `ifs_physics_common.config`, used at reference
`drivers/config.py:28-48` and passed to every component as `gt4py_config`)."""
from __future__ import annotations

from dataclasses import dataclass, field, replace
from typing import Any, Dict, Optional

import numpy as np


@dataclass(frozen=True)
class DataTypes:
    bool: type = bool
    float: type = np.float64
    int: type = np.int64


@dataclass
class GT4PyConfig:
    """Keeps the reference's field names.  `backend` is informational: there is exactly one
    backend, the hand-written sm_100a kernels ("b200")."""

    backend: str = "b200"
    backend_opts: Dict[str, Any] = field(default_factory=dict)
    build_info: Optional[Dict[str, Any]] = None
    device_sync: bool = True
    dtypes: DataTypes = field(default_factory=DataTypes)
    exec_info: Optional[Dict[str, Any]] = None
    managed: bool = False
    rebuild: bool = False
    validate_args: bool = False
    verbose: bool = True
    device: Optional[str] = None  # torch device string; None = "cuda" if available else "cpu"

    def with_backend(self, backend: Optional[str]) -> "GT4PyConfig":
        return replace(self, backend=backend or self.backend)

    def with_dtypes(self, dtypes: DataTypes) -> "GT4PyConfig":
        return replace(self, dtypes=dtypes)

    def with_validate_args(self, flag: bool) -> "GT4PyConfig":
        return replace(self, validate_args=flag)

    def reset_exec_info(self) -> None:
        if self.exec_info is not None:
            self.exec_info.clear()


@dataclass(frozen=True)
class GridConfig:
    nx: int
    ny: int
    nz: int


@dataclass
class IOConfig:
    output_csv_file: Optional[str] = None
    host_name: str = ""

    def with_output_csv_file(self, path: Optional[str]) -> "IOConfig":
        return replace(self, output_csv_file=path)

    def with_host_name(self, host_name: Optional[str]) -> "IOConfig":
        return replace(self, host_name=host_name or self.host_name)


@dataclass
class PythonConfig:
    """reference `drivers/config.py:28-47` (fluent `.with_*` builders, `run_nonlinear.py:210-217`)."""

    num_cols: Optional[int] = 1
    enable_validation: bool = True
    input_file: str = ""
    reference_file: str = ""
    num_runs: int = 1
    precision: str = "double"
    data_types: DataTypes = field(default_factory=DataTypes)
    gt4py_config: GT4PyConfig = field(default_factory=GT4PyConfig)
    sympl_enable_checks: bool = True
    num_threads: int = 1
    atol: float = 1e-16
    rtol: float = 1e-12

    def with_precision(self, precision: str) -> "PythonConfig":
        f = np.float64 if precision == "double" else np.float32
        i = np.int64 if precision == "double" else np.int32
        dtypes = DataTypes(bool=bool, float=f, int=i)
        return replace(self, precision=precision, data_types=dtypes, gt4py_config=self.gt4py_config.with_dtypes(dtypes))

    def with_backend(self, backend: Optional[str]) -> "PythonConfig":
        return replace(self, gt4py_config=self.gt4py_config.with_backend(backend))

    def with_checks(self, enabled: bool) -> "PythonConfig":
        return replace(self, sympl_enable_checks=enabled, gt4py_config=self.gt4py_config.with_validate_args(enabled))

    def with_validation(self, enabled: bool, atol: Optional[float] = None, rtol: Optional[float] = None) -> "PythonConfig":
        return replace(
            self, enable_validation=enabled, atol=self.atol if atol is None else atol, rtol=self.rtol if rtol is None else rtol
        )

    def with_num_cols(self, num_cols: Optional[int]) -> "PythonConfig":
        return replace(self, num_cols=num_cols if num_cols is not None else self.num_cols)

    def with_num_runs(self, num_runs: Optional[int]) -> "PythonConfig":
        return replace(self, num_runs=num_runs if num_runs is not None else self.num_runs)
