#!/usr/bin/env python
"""NL driver (reference: drivers/run_nonlinear.py:51-232): saturation + Cloudsc2NL, timed, validated.

    python -m drivers.run_nonlinear --num-cols 65536 --num-runs 20 --precision double

Validation: against the reference's golden file (`tests/golden/reference_<precision>.npz`) when the inputs come from a
matching `input.h5`; with synthetic inputs (the reference's input.h5 is not shipped) against the committed fixture
`tests/golden/oracle_base_<precision>.npz` (outputs for the first 16 columns of the seeded synthetic block)."""
from __future__ import annotations

import click
import numpy as np

from .config import DEFAULT_CONFIG, DEFAULT_IO_CONFIG, ROOT
from .common import FLOPS_PER_100_COLUMNS, problem, stats, write_performance_to_csv, write_stencils_performance_to_csv
from cloudsc2_b200.framework.stencil import resolve_exec_info
from cloudsc2_b200.framework.timing import Timer, timing
from cloudsc2_b200.physics.common.saturation import Saturation
from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL


def validate(fields, ref, atol, rtol, label):
    """Field-scaled comparison; prints like the reference's `validate`, returns all-ok."""
    ok = True
    for name, r in ref.items():
        got = fields[name]
        scale = max(np.abs(r).max(), np.finfo(np.float64).tiny)
        err = np.abs(got - r).max() / scale if np.abs(r).max() > 0 else np.abs(got).max()
        good = err <= max(rtol, atol / scale)
        ok &= bool(good)
        print(f"  {label}.{name:12s} {'OK ' if good else 'FAIL'} field-scaled max error {err:.3e}")
    return ok


def core(config, io_config):
    grid, state, dt, p, from_file = problem(config)
    cfg = config.gt4py_config
    nx = grid.nx
    saturation = Saturation(grid, kflag=1, lphylin=True, yoethf_params=p["yoethf"], yomcst_params=p["yomcst"],
                            enable_checks=config.sympl_enable_checks, gt4py_config=cfg)
    diags = saturation(state)
    state.update(diags)
    cloudsc2_nl = Cloudsc2NL(grid, lphylin=True, ldrain1d=False, yoethf_params=p["yoethf"], yomcst_params=p["yomcst"],
                             yrecldp_params=p["yrecldp"], yrephli_params=p["yrephli"], yrphnc_params=p["yrphnc"],
                             enable_checks=config.sympl_enable_checks, gt4py_config=cfg)
    tends, diags_cloudsc = cloudsc2_nl(state, dt)
    diags.update(diags_cloudsc)
    cfg.reset_exec_info()

    runtime_l = []
    for i in range(config.num_runs):
        Timer.reset()
        with timing(f"run_{i}"):
            saturation(state, out=diags)
            cloudsc2_nl(state, dt, out_tendencies=tends, out_diagnostics=diags)
        runtime_l.append(Timer.get_time(f"run_{i}", units="ms"))
    mean, std = stats(runtime_l)
    mflops = [FLOPS_PER_100_COLUMNS * (nx / 100) / (r * 1e-3) / 1e6 for r in runtime_l]
    mflops_mean, mflops_std = stats(mflops)
    print(f"Performance: {nx} columns, {mean:.3f} +/- {std:.3f} ms per run, {nx / mean / 1e3:.2f} M columns/s, "
          f"{mflops_mean:.0f} MFLOPS (reference flop count)")
    for name, rec in resolve_exec_info(cfg.exec_info).items():
        print(f"  {name}: {rec['ncalls']} calls, {rec['total_run_time'] * 1e3:.3f} ms")

    if io_config.output_csv_file is not None:
        write_performance_to_csv(io_config.output_csv_file, io_config.host_name, config.precision, "nl-" + cfg.backend, nx,
                                 config.num_threads, 1, config.num_runs, mean, std, mflops_mean, mflops_std)

    if config.enable_validation:
        print("\n== Validation:")
        out = {k: v.numpy() for k, v in {**tends, **diags}.items() if hasattr(v, "numpy")}
        if from_file:
            if config.reference_file.endswith(".h5"):  # a golden file in the reference's HDF5 form (data/reference_*.h5)
                from cloudsc2_b200.h5lite import File

                g = File(config.reference_file)
            else:
                g = np.load(config.reference_file)
            nz, klon = grid.nz, int(g["KLON"][0])
            cols = np.arange(nx) % klon
            pad = lambda a: np.vstack([a, np.zeros((1, a.shape[1]))]) if a.shape[0] == nz else a  # noqa: E731
            ref = {"f_clc": pad(g["PCLC"])[:, cols], "f_fplsn": g["PFPLSN"][:, cols], "f_fhpsn": g["PFHPSN"][:, cols],
                   "f_fplsl": g["PFPLSL"][:, cols], "f_fhpsl": g["PFHPSL"][:, cols], "f_t": pad(g["TENDENCY_LOC_T"])[:, cols],
                   "f_q": pad(g["TENDENCY_LOC_Q"])[:, cols], "f_ql": pad(g["TENDENCY_LOC_CLD"][0])[:, cols],
                   "f_qi": pad(g["TENDENCY_LOC_CLD"][1])[:, cols]}
            label = "golden"
        else:
            import os

            fx = np.load(os.path.join(ROOT, "tests", "golden", f"oracle_base_{config.precision}.npz"))
            n = min(nx, 16)
            ref = {k[5:]: fx[k][:, :n] for k in fx.files if k.startswith("nl_t_") or k.startswith("nl_d_")}
            out = {k: v[:, :n] for k, v in out.items()}
            label = "fixture"
        ok = validate(out, ref, config.atol, config.rtol, label)
        print("validation passed" if ok else "validation FAILED")
    from dataclasses import replace

    return replace(config, num_cols=nx)


@click.command()
@click.option("--backend", type=str, default=None, help="Ignored: the only backend is the hand-written sm_100a kernels.")
@click.option("--enable-checks/--disable-checks", is_flag=True, type=bool, default=False)
@click.option("--enable-validation/--disable-validation", is_flag=True, type=bool, default=True)
@click.option("--num-cols", type=int, default=None, help="Number of domain columns (default: 100).")
@click.option("--num-runs", type=int, default=1)
@click.option("--precision", type=click.Choice(("double", "single")), default="double")
@click.option("--host-alias", type=str, default=None)
@click.option("--output-csv-file", type=str, default=None)
@click.option("--output-csv-file-stencils", type=str, default=None, help="per-stencil device times (exec_info), one row per stencil")
@click.option("--reference-file", type=str, default=None, help="golden NL outputs for --input-file (.h5 like data/reference_*.h5, or .npz)")
@click.option("--input-file", type=str, default=None, help="input.h5 (default: tests/golden/input.h5 if present, else synthetic)")
@click.option("--atol", type=float, default=None)
@click.option("--rtol", type=float, default=None)
def main(backend, enable_checks, enable_validation, num_cols, num_runs, precision, host_alias, output_csv_file,
         output_csv_file_stencils, reference_file, input_file, atol, rtol):
    rtol = rtol if rtol is not None else (1e-12 if precision == "double" else 1e-5)
    config = (DEFAULT_CONFIG.with_precision(precision).with_backend(backend).with_checks(enable_checks)
              .with_validation(enable_validation, atol if atol is not None else 0.0, rtol).with_num_cols(num_cols or 0)
              .with_num_runs(num_runs))
    if input_file:
        config.input_file = input_file
    if reference_file:
        config.reference_file = reference_file
    config.gt4py_config.exec_info = {}
    io_config = DEFAULT_IO_CONFIG.with_output_csv_file(output_csv_file).with_host_name(host_alias)
    config = core(config, io_config)
    if output_csv_file_stencils is not None:  # reference drivers/run_nonlinear.py:221-232
        write_stencils_performance_to_csv(output_csv_file_stencils, io_config.host_name, config.precision,
                                          "nl-" + config.gt4py_config.backend, config.num_cols, config.num_threads,
                                          config.num_runs, config.gt4py_config.exec_info, key_patterns=["cloudsc", "saturation"])


if __name__ == "__main__":
    main()
