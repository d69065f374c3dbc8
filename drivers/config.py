"""Default driver configuration (reference: drivers/config.py:28-48)."""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, replace

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
for _p in (ROOT, os.path.join(ROOT, "gt4py-dwarf-p-cloudsc2-tl-ad_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from cloudsc2_b200.framework.config import DataTypes, GT4PyConfig, IOConfig, PythonConfig  # noqa: E402

DATA_DIR = os.path.join(ROOT, "tests", "golden")


@dataclass
class Config(PythonConfig):
    def with_precision(self, precision):
        cfg = super().with_precision(precision)
        return replace(cfg, reference_file=os.path.join(DATA_DIR, f"reference_{precision}.npz"))


DEFAULT_CONFIG = Config(
    num_cols=1,
    enable_validation=True,
    input_file=os.path.join(DATA_DIR, "input.h5"),  # used when present, synthetic inputs otherwise
    reference_file="",
    num_runs=1,
    precision="double",
    data_types=DataTypes(bool=bool, float=np.float64, int=np.int64),
    gt4py_config=GT4PyConfig(backend="b200", rebuild=False, validate_args=True, verbose=True),
    sympl_enable_checks=True,
)
DEFAULT_IO_CONFIG = IOConfig(output_csv_file=None, host_name="")
