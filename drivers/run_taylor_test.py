#!/usr/bin/env python
"""TL Taylor-test driver (reference: drivers/run_taylor_test.py:41-126).

    python -m drivers.run_taylor_test --num-cols 65536 --num-runs 5
"""
from __future__ import annotations

import click

from .config import DEFAULT_CONFIG, DEFAULT_IO_CONFIG
from .common import problem, stats, write_performance_to_csv
from cloudsc2_b200.framework.timing import Timer
from cloudsc2_b200.physics.tangent_linear.validation import TaylorTest


def core(config, io_config, fused=False):
    grid, state, dt, p, _ = problem(config)
    cfg = config.gt4py_config
    tt = TaylorTest(grid, factor1=0.01, factor2s=tuple(float(10 ** -(i + 1)) for i in range(0, 10)), kflag=1, lphylin=True,
                    ldrain1d=False, yoethf_params=p["yoethf"], yomcst_params=p["yomcst"], yrecldp_params=p["yrecldp"],
                    yrephli_params=p["yrephli"], yrncl_params=p["yrncl"], yrphnc_params=p["yrphnc"],
                    enable_checks=config.sympl_enable_checks, gt4py_config=cfg, fused=fused)
    norms = tt.run(state, dt)
    cfg.reset_exec_info()
    runtime_l, norms_l = [], []
    for _ in range(config.num_runs):
        Timer.reset()
        _ = tt.run(state, dt)
        runtime_l.append(Timer.get_time("run", units="ms"))
        norms_l.append(Timer.get_time("norms", units="ms"))
    mean, std = stats(runtime_l)
    passed, code = tt.validate(norms)
    print(f"\nThe test completed in {mean:.3f} ± {std:.3f} ms (+ {stats(norms_l)[0]:.3f} ms for the device-side norms).")
    if io_config.output_csv_file is not None:
        write_performance_to_csv(io_config.output_csv_file, io_config.host_name, config.precision, "tl-" + cfg.backend,
                                 grid.nx, config.num_threads, 1, config.num_runs, mean, std, 0, 0)
    return passed


@click.command()
@click.option("--enable-checks/--disable-checks", is_flag=True, type=bool, default=False)
@click.option("--num-cols", type=int, default=None)
@click.option("--num-runs", type=int, default=1)
@click.option("--precision", type=click.Choice(("double", "single")), default="double")
@click.option("--host-alias", type=str, default=None)
@click.option("--output-csv-file", type=str, default=None)
@click.option("--input-file", type=str, default=None)
@click.option("--fused/--unfused", is_flag=True, default=False, help="fuse PerturbedState into the NL kernel and StateIncrement into the TL kernel (same results)")
@click.option("--fused-sums", is_flag=True, default=False,
              help="one sweep per factor: increment, perturbation, NL and the field sums fused (same norms up to summation order)")
def main(enable_checks, num_cols, num_runs, precision, host_alias, output_csv_file, input_file, fused, fused_sums):
    config = (DEFAULT_CONFIG.with_precision(precision).with_checks(enable_checks).with_num_cols(num_cols or 100)
              .with_num_runs(num_runs))
    if input_file:
        config.input_file = input_file
    io_config = DEFAULT_IO_CONFIG.with_output_csv_file(output_csv_file).with_host_name(host_alias)
    raise SystemExit(0 if core(config, io_config, "sums" if fused_sums else fused) else 1)


if __name__ == "__main__":
    main()
