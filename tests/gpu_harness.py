"""Runs the product (components -> ctypes -> CUDA kernels) on a GPU and returns host copies of
every output in the oracle's `[nz+1, nx]` orientation.  Used by the `-m gpu` tests and smoke()."""
from __future__ import annotations

from datetime import timedelta
from typing import Any, Dict

import numpy as np

import helpers as H
from cloudsc2_b200 import iox, setup
from cloudsc2_b200.framework.config import DataTypes, GridConfig, GT4PyConfig
from cloudsc2_b200.framework.grid import ComputationalGrid
from cloudsc2_b200.framework.storage import Field
from cloudsc2_b200.physics.adjoint.microphysics import Cloudsc2AD
from cloudsc2_b200.physics.adjoint.validation import SymmetryTest
from cloudsc2_b200.physics.common.diagnostics import EtaLevels
from cloudsc2_b200.physics.common.increment import PerturbedState, StateIncrement
from cloudsc2_b200.physics.common.saturation import Saturation
from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL
from cloudsc2_b200.physics.tangent_linear.microphysics import Cloudsc2TL
from cloudsc2_b200.physics.tangent_linear.validation import TaylorTest


def to_host(d: Dict[str, Any]) -> Dict[str, np.ndarray]:
    return {k: v.numpy() for k, v in d.items() if isinstance(v, Field)}


def config_for(dtype) -> GT4PyConfig:
    i = np.int64 if np.dtype(dtype) == np.float64 else np.int32
    return GT4PyConfig(dtypes=DataTypes(bool=bool, float=np.dtype(dtype).type, int=i))


def make_grid_state(block: str, dtype, ncol: int, nz: int = 137):
    cfg = config_for(dtype)
    grid = ComputationalGrid(GridConfig(nx=ncol, ny=1, nz=nz))
    state = setup.get_synthetic_state(grid, gt4py_config=cfg, block=block)
    if ncol == 0:  # eta comes from GLOBAL column 0, which an empty shard does not own
        g1 = ComputationalGrid(GridConfig(nx=1, ny=1, nz=nz))
        s1 = setup.get_synthetic_state(g1, gt4py_config=cfg, block=block)
        state.update(EtaLevels(g1, gt4py_config=cfg)(s1))
    else:
        state.update(EtaLevels(grid, gt4py_config=cfg)(state))
    return cfg, grid, state


def run_components(block: str = "base", dtype=np.float64, ncol: int = 100, ad_predicates: str = "tl",
                   lregcl: bool = True, ad_trajectory: str = "checkpoint", nz: int = 137, dt_seconds: float = H.DT,
                   **flags) -> Dict[str, Any]:
    """saturation -> NL -> increment -> TL -> AD (symmetry pipeline) on the GPU."""
    cfg, grid, state = make_grid_state(block, dtype, ncol, nz)
    p = iox.ifs_defaults()
    p["yrncl"].LREGCL = lregcl
    p["yrphnc"].LEVAPLS2 = bool(flags.get("levapls2", False))
    lphylin, ldrain1d = bool(flags.get("lphylin", True)), bool(flags.get("ldrain1d", False))
    dt = timedelta(seconds=dt_seconds)
    out: Dict[str, Any] = {}

    sat = Saturation(grid, int(flags.get("kflag", 1)), lphylin, p["yoethf"], p["yomcst"], gt4py_config=cfg)
    state.update(sat(state))
    out["qsat"] = state["f_qsat"].numpy()
    out["eta"] = state["f_eta"].numpy()
    nl = Cloudsc2NL(grid, lphylin, ldrain1d, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrphnc"], gt4py_config=cfg)
    tn, dg = nl(state, dt)
    out["tends_nl"], out["diags_nl"] = to_host(tn), to_host(dg)
    if flags.get("nl_only"):
        return out

    st = SymmetryTest(grid, 0.01, int(flags.get("kflag", 1)), lphylin, ldrain1d, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"],
                      p["yrphnc"], gt4py_config=cfg, ad_predicates=ad_predicates, ad_trajectory=ad_trajectory,
                      fused=bool(flags.get("fused", False)))
    if "ignore_supsat" in flags:  # the symmetry harness ignores supsat; the Taylor harness does not
        st.state_increment = StateIncrement(grid, 0.01, ignore_supsat=bool(flags["ignore_supsat"]), gt4py_config=cfg)
    # run the pipeline step by step to keep host copies of the TL outputs before AD consumes them
    st.diags_sat = st.saturation(state, out=st.diags_sat)
    state.update(st.diags_sat)
    st.state_i = st.state_increment(state, out=st.state_i)
    state.update(st.state_i)
    out["state_i"] = to_host(st.state_i)
    st.tends_tl, st.diags_tl = st.cloudsc2_tl(state, dt, out_tendencies=st.tends_tl, out_diagnostics=st.diags_tl)
    out["tends_tl"], out["diags_tl"] = to_host(st.tends_tl), to_host(st.diags_tl)
    if flags.get("tl_only"):
        return out
    norm1 = st.get_norm1(st.tends_tl, st.diags_tl)
    st.add_tendencies_to_state(state, st.tends_tl)
    state.update(st.diags_tl)
    st.tends_ad, st.diags_ad = st.cloudsc2_ad(state, dt, out_tendencies=st.tends_ad, out_diagnostics=st.diags_ad)
    out["tends_ad"], out["diags_ad"] = to_host(st.tends_ad), to_host(st.diags_ad)
    out["seeds_after"] = {k: state[k].numpy() for k in
                          ("f_tnd_t_i", "f_tnd_q_i", "f_tnd_ql_i", "f_tnd_qi_i", "f_clc_i", "f_covptot_i", "f_fhpsl_i",
                           "f_fhpsn_i", "f_fplsl_i", "f_fplsn_i")}
    norm2 = st.get_norm2(st.state_i, st.tends_ad, st.diags_ad)
    eps = float(np.finfo(np.dtype(dtype)).eps)
    n1, n2 = norm1.cpu().numpy(), norm2.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        n3 = np.where(n2 == 0, abs(n1 - n2) / eps, abs(n1 - n2) / (eps * n2))
    out["norm1"], out["norm2"], out["norm3"] = n1, n2, n3
    out["symmetry_norm3_max"] = float(n3.max())
    return out


def run_taylor(block: str = "base", dtype=np.float64, ncol: int = 100, fused=False):
    cfg, grid, state = make_grid_state(block, dtype, ncol)
    p = iox.ifs_defaults()
    tt = TaylorTest(grid, 0.01, tuple(float(10 ** -(i + 1)) for i in range(10)), 1, True, False, p["yoethf"], p["yomcst"],
                    p["yrecldp"], p["yrephli"], p["yrncl"], p["yrphnc"], gt4py_config=cfg, fused=fused)
    norms = tt.run(state, timedelta(seconds=H.DT))
    return tt, norms


def run_symmetry(block: str = "base", dtype=np.float64, ncol: int = 100, ad_predicates: str = "tl", fused: bool = False):
    cfg, grid, state = make_grid_state(block, dtype, ncol)
    p = iox.ifs_defaults()
    st = SymmetryTest(grid, 0.01, 1, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"],
                      p["yrphnc"], gt4py_config=cfg, ad_predicates=ad_predicates, fused=fused)
    passed = st(state, timedelta(seconds=H.DT), verbose=False)
    return st, passed
