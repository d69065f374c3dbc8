#!/usr/bin/env python
"""AD symmetry-test driver (reference: drivers/run_symmetry_test.py:41-123).

    python -m drivers.run_symmetry_test --num-cols 65536 --num-runs 5 [--ad-predicates reference]
"""
from __future__ import annotations

import click

from .config import DEFAULT_CONFIG, DEFAULT_IO_CONFIG
from .common import problem, stats, write_performance_to_csv
from cloudsc2_b200.framework.timing import Timer, timing
from cloudsc2_b200.physics.adjoint.validation import SymmetryTest


def core(config, io_config, ad_predicates=None, fused=False, synthetic_block="cold"):
    grid, state, dt, p, from_file = problem(config, block=synthetic_block)
    if not from_file:
        print(f"no input file: synthetic '{synthetic_block}' block (the reference's shipped input is all-cold; its literal AD predicates "
              f"fail its own symmetry test on columns that cross RTT inside a level, i.e. on the 'base' block)")
    cfg = config.gt4py_config
    st = SymmetryTest(grid, factor=0.01, kflag=1, lphylin=True, ldrain1d=False, yoethf_params=p["yoethf"],
                      yomcst_params=p["yomcst"], yrecldp_params=p["yrecldp"], yrephli_params=p["yrephli"],
                      yrncl_params=p["yrncl"], yrphnc_params=p["yrphnc"], enable_checks=config.sympl_enable_checks,
                      gt4py_config=cfg, ad_predicates=ad_predicates, fused=fused)
    print(f"AD branch predicates: {st.cloudsc2_ad.ad_predicates!r} ('reference' = the reference AD stencil literally; "
          f"'tl' = the TL sweep's predicates, an exact adjoint on any input; --ad-predicates / CS2_AD_PREDICATES)")
    passed = st(state, dt, enable_validation=True)
    cfg.reset_exec_info()
    runtime_l = []
    for i in range(config.num_runs):
        Timer.reset()
        with timing(f"run_{i}"):
            st(state, dt, enable_validation=False)
        runtime_l.append(Timer.get_time(f"run_{i}", units="ms"))
    mean, std = stats(runtime_l)
    print(f"\nThe test completed in {mean:.3f} ± {std:.3f} ms.")
    if io_config.output_csv_file is not None:
        write_performance_to_csv(io_config.output_csv_file, io_config.host_name, config.precision, "ad-" + cfg.backend,
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
@click.option("--ad-predicates", type=click.Choice(("tl", "reference")), default=None)
@click.option("--fused/--unfused", is_flag=True, default=False, help="form the TL perturbation inside the TL kernel (same results)")
@click.option("--synthetic-block", type=click.Choice(("base", "cold")), default="cold", help="synthetic stand-in for a missing input file")
def main(enable_checks, num_cols, num_runs, precision, host_alias, output_csv_file, input_file, ad_predicates, fused, synthetic_block):
    config = (DEFAULT_CONFIG.with_precision(precision).with_checks(enable_checks).with_num_cols(num_cols or 100)
              .with_num_runs(num_runs))
    if input_file:
        config.input_file = input_file
    io_config = DEFAULT_IO_CONFIG.with_output_csv_file(output_csv_file).with_host_name(host_alias)
    raise SystemExit(0 if core(config, io_config, ad_predicates, fused, synthetic_block) else 1)


if __name__ == "__main__":
    main()
