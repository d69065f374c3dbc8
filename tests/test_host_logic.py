"""CPU: host-side mirror of the reference interface (storage layout, params, components,
harness scoring, sharding, the gloo world_size-2 path)."""
import os
import subprocess
import sys
from datetime import timedelta

import numpy as np
import pytest
import torch

import helpers as H
from cloudsc2_b200 import distributed, iox, setup, synthetic
from cloudsc2_b200._lib import CUDAExtensionError
from cloudsc2_b200.framework.config import GridConfig, GT4PyConfig, PythonConfig
from cloudsc2_b200.framework.grid import ComputationalGrid, I, J, K
from cloudsc2_b200.framework.storage import zeros
from cloudsc2_b200.physics.adjoint.microphysics import Cloudsc2AD
from cloudsc2_b200.physics.common.diagnostics import EtaLevels
from cloudsc2_b200.physics.common.increment import PerturbedState, StateIncrement
from cloudsc2_b200.physics.common.saturation import Saturation
from cloudsc2_b200.physics.nonlinear.microphysics import Cloudsc2NL
from cloudsc2_b200.physics.tangent_linear.microphysics import Cloudsc2TL
from cloudsc2_b200.physics.tangent_linear.validation import TaylorTest

CPU = GT4PyConfig(device="cpu")


def test_storage_layout_and_logical_view():
    grid = ComputationalGrid(GridConfig(nx=100, ny=1, nz=137))
    f = zeros(grid, (I, J, K), gt4py_config=CPU, units="K", name="f_t")
    assert f.buffer.shape == (138, 128) and f.buffer.stride() == (128, 1)
    assert f.data.shape == (100, 1, 138) and f.data.stride()[0] == 1 and f.data.stride()[2] == 128
    assert f.data.data_ptr() == f.buffer.data_ptr()  # zero-copy view
    arr = np.arange(137 * 100, dtype=np.float64).reshape(137, 100)
    f.assign(arr)
    assert np.array_equal(f.numpy()[:137], arr) and not f.numpy()[137].any()
    assert float(f.data[7, 0, 3]) == arr[3, 7]
    h = zeros(grid, (I, J, K - 1 / 2), gt4py_config=CPU)
    assert h.data.shape == (100, 1, 138)
    # one view object per (buffer, nx): handed out again, rebuilt when either changes
    view = f.data
    assert f.data is view
    f.buffer = f.buffer.clone()
    assert f.data is not view and f.data.data_ptr() == f.buffer.data_ptr()
    view = f.data
    f.nx = 64
    assert f.data is not view and f.data.shape == (64, 1, 138)
    assert grid.grids[I, J, K].shape == (100, 1, 137) and grid.grids[I, J, K - 1 / 2].shape == (100, 1, 138)


def test_params_models_behave_like_the_reference():
    d = iox.ifs_defaults()
    ext = {}
    for v in d.values():
        ext.update(v.dict())
    for name in ("R2ES", "R5ALVCP", "RLSTT", "RKCONV", "RLPTRC", "LREGCL", "LEVAPLS2", "RVTMP2"):
        assert name in ext
    assert d["yrncl"].LREGCL is True and d["yrphnc"].LEVAPLS2 is False and ext["RVTMP2"] == 0.0
    d["yrncl"].LREGCL = False  # in-place mutation used by the Taylor harness
    assert d["yrncl"].dict()["LREGCL"] is False
    with pytest.raises(TypeError):
        iox.YomcstParams(RCPD=1.0)
    extra = iox.YrecldpParams(RCLCRIT=1.0, RKCONV=2.0, RLMIN=3.0, RPECONS=4.0, NCLDTOP=15)
    assert extra.dict()["NCLDTOP"] == 15
    assert abs(ext["RLSTT"] - 2.8345e6) == 0 and abs(ext["R5LES"] - ext["R3LES"] * (ext["RTT"] - ext["R4LES"])) == 0


def test_component_surface_and_loud_failure_without_gpu():
    grid = ComputationalGrid(GridConfig(nx=64, ny=1, nz=137))
    state = setup.get_synthetic_state(grid, gt4py_config=CPU)
    assert set(state) >= {"f_ap", "f_aph", "f_t", "f_q", "f_ql", "f_qi", "f_lu", "f_lude", "f_mfu", "f_mfd", "f_supsat",
                          "f_tnd_cml_t", "f_tnd_cml_q", "f_tnd_cml_ql", "f_tnd_cml_qi", "time"}
    eta = EtaLevels(grid, gt4py_config=CPU)(state)
    state.update(eta)
    ref_eta = H.onp.eta_levels(state["f_ap"].numpy(), state["f_aph"].numpy())
    assert np.array_equal(state["f_eta"].numpy(), ref_eta)
    p = iox.ifs_defaults()
    sat = Saturation(grid, 1, True, p["yoethf"], p["yomcst"], gt4py_config=CPU)
    nl = Cloudsc2NL(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrphnc"], gt4py_config=CPU)
    tl = Cloudsc2TL(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"], p["yrphnc"], gt4py_config=CPU)
    ad = Cloudsc2AD(grid, True, False, p["yoethf"], p["yomcst"], p["yrecldp"], p["yrephli"], p["yrncl"], p["yrphnc"], gt4py_config=CPU)
    assert set(nl.tendency_grid_properties) == {"f_q", "f_qi", "f_ql", "f_t"}
    assert set(nl.diagnostic_grid_properties) == {"f_clc", "f_covptot", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn"}
    assert len(nl.input_grid_properties) == 17 and len(tl.input_grid_properties) == 33
    assert len(tl.tendency_grid_properties) == 8 and len(tl.diagnostic_grid_properties) == 12
    assert set(ad.tendency_grid_properties) == {"f_t", "f_q", "f_ql", "f_qi", "f_cml_t_i", "f_cml_q_i", "f_cml_ql_i", "f_cml_qi_i"}
    assert len(ad.diagnostic_grid_properties) == 18 and len(ad.input_grid_properties) == 27
    assert len(StateIncrement(grid, 0.01, gt4py_config=CPU).diagnostic_grid_properties) == 16
    assert len(PerturbedState(grid, 0.1, gt4py_config=CPU).input_grid_properties) == 32
    if not torch.cuda.is_available():
        with pytest.raises(CUDAExtensionError, match="no CPU fallback"):
            sat(state)
        state["f_qsat"] = zeros(grid, (I, J, K), gt4py_config=CPU)
        with pytest.raises(CUDAExtensionError):
            nl(state, timedelta(seconds=3600))
    with pytest.raises(KeyError, match="f_qsat|missing"):
        nl({k: v for k, v in state.items() if k != "f_qsat"}, timedelta(seconds=3600))


def test_taylor_scoring_matches_reference_rules():
    grid = ComputationalGrid(GridConfig(nx=32, ny=1, nz=137))
    p = iox.ifs_defaults()
    tt = TaylorTest(grid, 0.01, tuple(10.0 ** -(i + 1) for i in range(10)), 1, True, False, p["yoethf"], p["yomcst"],
                    p["yrecldp"], p["yrephli"], p["yrncl"], p["yrphnc"], gt4py_config=CPU)
    assert p["yrncl"].LREGCL is False  # the harness switches the regularisation off
    good = 1 + np.array([1e-1, 1e-2, 1e-3, 1e-4, 1e-5, 1e-6, 1e-7, 1e-8, 1e-6, 1e-4])
    assert tt.validate(good, verbose=False) == (True, 0)
    assert tt.validate(np.full(10, 2.0), verbose=False) == (False, 13)
    flat = 1 + np.array([1e-1, 1e-2, 1e-3, 1e-4, 2e-5, 3e-5, 4e-5, 5e-5, 6e-5, 7e-5])
    assert tt.validate(flat, verbose=False) == (False, 12)
    assert H.onp.taylor_score(good)[:2] == (True, 0) and H.onp.taylor_score(flat)[:2] == (False, 12)


def test_column_sharding_covers_everything_once():
    for n, w in ((1_048_576, 8), (100, 3), (7, 8), (65_536, 1)):
        blocks = [distributed.shard_columns(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= 1


def test_synthetic_block_is_two_sided():
    blk = synthetic.base_block()
    t = blk["f_t"][:137]
    assert (t < 273.16).any() and (t > 275.16).any() and (t < 250.16).any()
    assert (blk["f_lude"] >= 1e-7).any() and (blk["f_lu"] >= 1e-10).any()
    tiled = synthetic.tile(blk, 250)
    assert np.array_equal(tiled["f_t"][:, 100:200], blk["f_t"]) and tiled["f_t"].shape == (138, 250)
    assert synthetic.cold_block()["f_t"].max() < 273.16


def test_python_config_builders():
    c = PythonConfig().with_precision("single").with_num_cols(4096).with_num_runs(5).with_validation(True, 1e-9, 1e-6)
    assert c.gt4py_config.dtypes.float is np.float32 and c.num_cols == 4096 and c.num_runs == 5 and c.rtol == 1e-6


def test_world_size_2_gloo_reductions():
    """The N>1 host path (shard -> local sums -> all-reduce SUM / MAX, eta broadcast) on CPU."""
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dist_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29641")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29641", script],
        env=env, capture_output=True, text=True, timeout=300,
    )
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "DIST_OK rank=0" in res.stdout and "DIST_OK rank=1" in res.stdout


def test_product_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU arms may touch oracle/ (or tests/helpers.py, which
    imports it): the package and the drivers must not."""
    import re

    offenders = []
    for top in ("gt4py-dwarf-p-cloudsc2-tl-ad_b200", "drivers"):
        for dirpath, _, files in os.walk(os.path.join(H.ROOT, top)):
            for fn in files:
                if fn.endswith(".py"):
                    text = open(os.path.join(dirpath, fn)).read()
                    if re.search(r"^\s*(import|from)\s+(oracle|helpers|gpu_harness)\b", text, flags=re.M):
                        offenders.append(os.path.join(dirpath, fn))
    assert not offenders, offenders
