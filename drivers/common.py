"""Shared driver plumbing: grid/state/parameters from `input.h5` when available, else synthetic."""
from __future__ import annotations

import csv
import os
from datetime import timedelta
from typing import Any, Dict, List, Tuple

from .config import Config  # noqa: F401  (sets sys.path)
from cloudsc2_b200 import iox, setup, synthetic
from cloudsc2_b200.framework.config import GridConfig
from cloudsc2_b200.framework.grid import ComputationalGrid
from cloudsc2_b200.physics.common.diagnostics import EtaLevels

FLOPS_PER_100_COLUMNS = 12482329  # hard-coded CLOUDSC HPM count behind the reference's "MFLOPS" 


def problem(config, block: str = "base") -> Tuple[ComputationalGrid, Dict[str, Any], timedelta, Dict[str, Any], bool]:
    """(grid, state incl. f_eta, timestep, parameter sets, from_file).  `block`: which seeded synthetic block stands in
    for a missing input file ("base": warm and cold columns; "cold": all-cold columns like the reference's shipped input)."""
    cfg = config.gt4py_config
    from_file = bool(config.input_file) and os.path.exists(config.input_file)
    if from_file:
        op = iox.HDF5Operator(config.input_file, gt4py_config=cfg)
        nx = config.num_cols or op.get_nlon()
        grid = ComputationalGrid(GridConfig(nx=nx, ny=1, nz=op.get_nlev()))
        state = setup.get_state(setup.HDF5GridOperator(config.input_file, grid, gt4py_config=cfg))
        dt = op.get_timestep()
        params = {
            "yoethf": op.get_yoethf_params(), "yomcst": op.get_yomcst_params(), "yrecldp": op.get_yrecldp_params(),
            "yrephli": op.get_yrephli_params(), "yrncl": op.get_yrncl_params(), "yrphnc": op.get_yrphnc_params(),
        }
    else:
        nx = config.num_cols or synthetic.KLON
        grid = ComputationalGrid(GridConfig(nx=nx, ny=1, nz=synthetic.KLEV))
        state = setup.get_synthetic_state(grid, gt4py_config=cfg, block=block)
        dt = iox.DEFAULT_TIMESTEP
        params = iox.ifs_defaults()
    state.update(EtaLevels(grid, enable_checks=config.sympl_enable_checks, gt4py_config=cfg)(state))
    return grid, state, dt, params, from_file


def stats(runtimes_ms: List[float]) -> Tuple[float, float]:
    n = len(runtimes_ms)
    mean = sum(runtimes_ms) / n
    std = (sum((r - mean) ** 2 for r in runtimes_ms) / (n - 1 if n > 1 else n)) ** 0.5
    return mean, std


def write_performance_to_csv(path, host_name, precision, variant, num_cols, num_threads, nproma, num_runs, runtime_mean,
                             runtime_stddev, mflops_mean, mflops_stddev) -> None:
    """One CSV row with the reference's columns (drivers/run_nonlinear.py:124-137)."""
    new = not os.path.exists(path)
    with open(path, "a", newline="") as fh:
        w = csv.writer(fh, delimiter=",")
        if new:
            w.writerow(["date", "host", "precision", "variant", "num_cols", "num_threads", "nproma", "num_runs",
                        "runtime_mean", "runtime_stddev", "mflops_mean", "mflops_stddev"])
        import datetime

        w.writerow([datetime.date.today().strftime("%Y%m%d"), host_name, precision, variant, num_cols, num_threads, nproma,
                    num_runs, runtime_mean, runtime_stddev, mflops_mean, mflops_stddev])


def write_stencils_performance_to_csv(path, host_name, precision, variant, num_cols, num_threads, num_runs, exec_info,
                                      key_patterns=("cloudsc", "saturation")) -> None:
    """Per-stencil timings (reference drivers/run_nonlinear.py:221-232 -> ifs_physics_common.output.
    write_stencils_performance_to_csv, whose exact column set is not in the reference tree): one row per stencil whose
    name contains one of `key_patterns`, with the number of calls and the device time GT4Py would report as
    `exec_info[<stencil>]["total_run_time"]` (here: CUDA events around the kernel launches of the stencil)."""
    from cloudsc2_b200.framework.stencil import resolve_exec_info

    import datetime

    new = not os.path.exists(path)
    with open(path, "a", newline="") as fh:
        w = csv.writer(fh, delimiter=",")
        if new:
            w.writerow(["date", "host", "precision", "variant", "num_cols", "num_threads", "num_runs", "stencil", "ncalls",
                        "total_run_time_ms", "run_time_per_call_ms"])
        for name, rec in sorted(resolve_exec_info(exec_info).items()):
            if not any(pat in name for pat in key_patterns):
                continue
            total_ms = rec["total_run_time"] * 1e3
            w.writerow([datetime.date.today().strftime("%Y%m%d"), host_name, precision, variant, num_cols, num_threads,
                        num_runs, name, rec["ncalls"], total_ms, total_ms / max(rec["ncalls"], 1)])
