"""CPU: the `input.h5` path -- h5lite reader/writer, iox.HDF5Operator, setup.HDF5GridOperator /
get_state, the golden loaders -- on an `input.h5` look-alike written with the reference's dataset names
(reference setup.py:28-70, iox.py:212-244, nonlinear/reference.py:28-55).  The reference's own data/input.h5 is not shipped."""
import os

import numpy as np
import pytest

import helpers as H
from cloudsc2_b200 import h5lite, iox, setup, synthetic
from cloudsc2_b200.framework.config import DataTypes, GridConfig, GT4PyConfig
from cloudsc2_b200.framework.grid import ComputationalGrid
from cloudsc2_b200.physics.nonlinear.reference import get_reference_diagnostics, get_reference_tendencies

CPU = GT4PyConfig(dtypes=DataTypes(bool=bool, float=np.float64, int=np.int64), device="cpu")


def test_h5lite_round_trip_many_datasets_and_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    data = {f"DS_{i:03d}": rng.normal(size=(3, 5)) for i in range(300)}  # 300 names: three symbol-table nodes
    data.update(A_F32=rng.normal(size=(7,)).astype(np.float32), B_I32=np.arange(6, dtype=np.int32).reshape(2, 3),
                C_I64=np.array([2**40, -3], dtype=np.int64), D_3D=rng.normal(size=(5, 4, 3)), E_SCALAR=np.float64(2.5),
                F_BOOL=np.array([True]), Z_EMPTY=np.zeros((0, 4)))
    path = str(tmp_path / "many.h5")
    h5lite.write_file(path, data)
    f = h5lite.File(path)
    assert set(f.keys()) == set(data)
    for k, v in data.items():
        got = f[k]
        exp = np.asarray(v).reshape(1) if np.ndim(v) == 0 else np.asarray(v)
        if exp.dtype == np.bool_:
            exp = exp.astype(np.int32)
        assert got.dtype == exp.dtype and got.shape == exp.shape and np.array_equal(got, exp), k
    with pytest.raises(NotImplementedError):
        h5lite.write_file(path, {"X": np.array(["a"])})


def test_reference_golden_files_and_written_files_share_the_reader():
    ref = "/root/reference/data/reference_double.h5"
    if not os.path.exists(ref):
        pytest.skip("reference data not on this box")
    g = h5lite.File(ref)
    fx = np.load(os.path.join(H.ROOT, "tests", "golden", "reference_double.npz"))
    for k in fx.files:
        assert np.array_equal(g[k], fx[k]), k


@pytest.fixture(scope="module")
def input_file(tmp_path_factory):
    path = str(tmp_path_factory.mktemp("h5") / "input.h5")
    synthetic.write_input_h5(path, block="base")
    return path


def test_hdf5_operator_reads_every_parameter_model(input_file):
    op = iox.HDF5Operator(input_file, gt4py_config=CPU)
    d = iox.ifs_defaults()
    assert op.get_nlon() == 100 and op.get_nlev() == 137 and op.get_timestep().total_seconds() == 3600.0
    assert op.get_yoethf_params().dict() == d["yoethf"].dict()
    assert op.get_yomcst_params().dict() == d["yomcst"].dict()
    assert op.get_yrecldp_params().dict() == d["yrecldp"].dict()   # datasets prefixed YRECLDP_ (iox.py:232)
    assert op.get_yrephli_params().dict() == d["yrephli"].dict()   # datasets prefixed YREPHLI_ (iox.py:237)
    assert op.get_yrncl_params().LREGCL is True and op.get_yrphnc_params().LEVAPLS2 is False  # defaults (iox.py:205,209)
    assert "YRECLDP_RCLCRIT" in op.f and "YREPHLI_RLPTRC" in op.f and "RCLCRIT" not in op.f


@pytest.mark.parametrize("nx,offset", [(100, 0), (250, 0), (64, 37)])
def test_get_state_tiles_columns_like_num_cols(input_file, nx, offset):
    """column i of the grid <- column (i + offset) mod KLON of the file (setup.py:66-68; --num-cols > KLON replicates)."""
    grid = ComputationalGrid(GridConfig(nx=nx, ny=1, nz=137))
    state = setup.get_state(setup.HDF5GridOperator(input_file, grid, gt4py_config=CPU, column_offset=offset))
    blk = synthetic.base_block()
    cols = (np.arange(nx) + offset) % 100
    assert set(state) == set(setup.FIELD_PROPERTIES) | {"time"}
    for name in setup.FIELD_PROPERTIES:
        got = state[name].numpy()
        assert got.shape == (138, nx)
        if name == "f_a":
            assert not got.any()
            continue
        rows = 138 if name == "f_aph" else 137
        assert np.array_equal(got[:rows], blk[name][:rows, cols]), name
        assert not got[rows:].any()  # the padding level of full-level fields stays zero
    assert state["f_aph"].data.shape == (nx, 1, 138)  # the logical (nx, 1, nz+1) view the harnesses index


def test_get_state_raises_on_a_missing_dataset(tmp_path):
    d = synthetic.input_h5_datasets(synthetic.base_block(), iox.ifs_defaults())
    del d["PLUDE"]
    path = str(tmp_path / "broken.h5")
    h5lite.write_file(path, d)
    grid = ComputationalGrid(GridConfig(nx=10, ny=1, nz=137))
    with pytest.raises(KeyError):
        setup.get_state(setup.HDF5GridOperator(path, grid, gt4py_config=CPU))
    with pytest.raises(KeyError):  # a required scalar
        d2 = synthetic.input_h5_datasets(synthetic.base_block(), iox.ifs_defaults())
        del d2["RTT"]
        h5lite.write_file(path, d2)
        iox.HDF5Operator(path).get_yomcst_params()


def test_golden_loaders_on_a_written_reference_file(tmp_path):
    P = H.externals()
    s = H.with_diagnostics(H.make_state("cold"), P)
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    path = str(tmp_path / "reference_double.h5")
    synthetic.write_reference_h5(path, tn, dg)
    grid = ComputationalGrid(GridConfig(nx=130, ny=1, nz=137))
    op = setup.HDF5GridOperator(path, grid, gt4py_config=CPU)
    tends, diags = get_reference_tendencies(op), get_reference_diagnostics(op)
    cols = np.arange(130) % 100
    for k in ("f_t", "f_q", "f_ql", "f_qi"):
        assert np.array_equal(tends[k].numpy(), tn[k][:, cols]), k
    assert tends["f_qv"] is tends["f_q"]
    for k in ("f_clc", "f_covptot", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn"):
        assert np.array_equal(diags[k].numpy(), dg[k][:, cols]), k
