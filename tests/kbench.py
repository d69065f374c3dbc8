#!/usr/bin/env python
"""Developer micro-benchmark (test infrastructure): kernel-only times of NL / TL / AD (CUDA events) + a quick parity
check against the oracle.
    python tests/kbench.py [--columns 65536] [--precision double] [--reps 20]
"""
import argparse
import json
import os
import sys
from datetime import timedelta

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path[:0] = [ROOT, os.path.join(ROOT, "gt4py-dwarf-p-cloudsc2-tl-ad_b200"), os.path.join(ROOT, "tests")]

import numpy as np
import torch

import gpu_harness as G
import helpers as H
from cloudsc2_b200 import iox
from cloudsc2_b200.physics.adjoint.validation import SymmetryTest
from cloudsc2_b200.physics.common.saturation import Saturation
from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL

ELEMS = {"nl": 3567, "tl": 7134, "ad": 8508, "ad_ckpt": 8508, "sat": 411}


def time_call(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--columns", type=int, nargs="+", default=[65536])
    ap.add_argument("--precision", default="double")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--fused", action="store_true")
    args = ap.parse_args()
    dtype = np.float64 if args.precision == "double" else np.float32
    es = np.dtype(dtype).itemsize
    if not args.no_parity:
        out = G.run_components(block="base", dtype=dtype, ncol=100)
        P = H.externals(LREGCL=True)
        _, _, _, ref = H.oracle_symmetry(H.make_state("base", dtype), P, predicates="tl")
        tn, dg = H.onp.cloudsc2_nl(ref["state"], H.DT, P)
        worst = 0.0
        for got, rf in ((out["tends_nl"], tn), (out["diags_nl"], dg), (out["tends_tl"], ref["tends_tl"]),
                        (out["diags_tl"], ref["diags_tl"]), (out["tends_ad"], ref["tends_ad"]), (out["diags_ad"], ref["diags_ad"])):
            for k, v in rf.items():
                if np.abs(v).max() > 0:
                    worst = max(worst, H.field_err(got[k], v))
        print(f"parity: worst field-scaled error vs oracle {worst:.3e}; symmetry {out['symmetry_norm3_max']:.1f} eps", flush=True)
    for ncol in args.columns:
        cfg, grid, state = G.make_grid_state("base", dtype, ncol)
        p = iox.ifs_defaults()
        dt = timedelta(seconds=3600)
        sat = Saturation(grid, 1, True, p["yoethf"], p["yomcst"], gt4py_config=cfg)
        state.update(sat(state))
        nl = Cloudsc2NL(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrphnc"], gt4py_config=cfg)
        tn, dg = nl(state, dt)
        st = SymmetryTest(grid, 0.01, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"], p["yrphnc"], gt4py_config=cfg, ad_trajectory="recompute")
        st(state, dt, enable_validation=False)
        st_ck = SymmetryTest(grid, 0.01, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"], p["yrphnc"], gt4py_config=cfg, ad_trajectory="checkpoint")
        st_ck.tends_tl, st_ck.diags_tl, st_ck.tends_ad, st_ck.diags_ad = st.tends_tl, st.diags_tl, st.tends_ad, st.diags_ad
        res = {"columns": ncol, "precision": args.precision}
        for name, fn in (
            ("sat", lambda: sat(state, out={"f_qsat": state["f_qsat"]})),
            ("nl", lambda: nl(state, dt, out_tendencies=tn, out_diagnostics=dg)),
            ("tl", lambda: st.cloudsc2_tl(state, dt, out_tendencies=st.tends_tl, out_diagnostics=st.diags_tl)),
            ("ad", lambda: st.cloudsc2_ad(state, dt, out_tendencies=st.tends_ad, out_diagnostics=st.diags_ad)),
            ("ad_ckpt", lambda: st_ck.cloudsc2_ad(state, dt, out_tendencies=st.tends_ad, out_diagnostics=st.diags_ad)),
        ):
            ms = time_call(fn, args.reps)
            gbs = ELEMS[name] * es * ncol / ms / 1e6
            res[name] = {"ms": round(ms, 4), "Mcol_s": round(ncol / ms / 1e3, 2), "GBs": round(gbs, 1), "frac": round(gbs / 6541.8, 4)}
        if args.fused:  # opt-in fused path: state_increment + TL in one sweep
            from cloudsc2_b200.physics.tangent_linear.microphysics import IncrementedCloudsc2TL

            tli = IncrementedCloudsc2TL(grid, 0.01, True, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"],
                                        p["yrncl"], p["yrphnc"], gt4py_config=cfg)
            for name, fn in (
                ("increment+tl (2 launches)", lambda: (st.state_increment(state, out=st.state_i),
                                                       st.cloudsc2_tl(state, dt, out_tendencies=st.tends_tl,
                                                                      out_diagnostics=st.diags_tl))),
                ("tl_increment fused", lambda: tli(state, dt, out_tendencies=st.tends_tl, out_diagnostics=st.diags_tl)),
            ):
                res[name] = {"ms": round(time_call(fn, args.reps), 4)}
        print(json.dumps(res), flush=True)
        del st, st_ck, nl, sat, state, tn, dg
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
