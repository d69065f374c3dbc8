"""CPU: the reference's golden NL outputs (data/reference_{double,single}.h5, committed loss-free as
tests/golden/reference_*.npz by tests/golden/make_golden.py).

`data/input.h5` is not shipped with the reference, so the golden outputs cannot be reproduced
point-wise here; what can be pinned are the structural invariants of the files and, as soon as an
`input.h5` is dropped at tests/golden/input.h5 (or $CS2_INPUT_H5), the full comparison."""
import os

import numpy as np
import pytest

import helpers as H
from cloudsc2_b200.h5lite import File

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
INPUT_H5 = os.environ.get("CS2_INPUT_H5", os.path.join(GOLDEN, "input.h5"))


@pytest.mark.parametrize("precision", ["double", "single"])
def test_golden_invariants(precision):
    g = np.load(os.path.join(GOLDEN, f"reference_{precision}.npz"))
    assert int(g["KLON"][0]) == 100 and int(g["KLEV"][0]) == 137
    assert g["PCLC"].shape == (137, 100) and g["PFPLSN"].shape == (138, 100)
    assert g["TENDENCY_LOC_CLD"].shape == (5, 137, 100)
    assert g["PCLC"].min() >= 0 and g["PCLC"].max() <= 1
    assert not g["PCOVPTOT"].any()
    assert not g["PFPLSN"][0].any() and not g["PFHPSN"][0].any()
    assert not g["TENDENCY_LOC_CLD"][2:].any()
    P = H.externals()
    rtol = 0 if precision == "double" else 1e-6
    # fhpsn = -RLSTT * fplsn (nonlinear/_stencils/cloudsc2.py:399) pins RLSTT = 2.8345e6
    np.testing.assert_allclose(g["PFHPSN"], -P["RLSTT"] * g["PFPLSN"], rtol=max(rtol, 2e-16), atol=0)
    if precision == "double":
        assert not g["PFPLSL"].any() and not g["PFHPSL"].any()  # all-cold input: no rain


def test_golden_single_vs_double_consistent():
    d = np.load(os.path.join(GOLDEN, "reference_double.npz"))
    s = np.load(os.path.join(GOLDEN, "reference_single.npz"))
    for name in ("PCLC", "PFPLSN", "PFHPSN", "TENDENCY_LOC_T", "TENDENCY_LOC_Q", "TENDENCY_LOC_CLD"):
        scale = np.abs(d[name]).max()
        assert np.abs(d[name] - s[name]).max() / scale < 1e-4, name


@pytest.mark.skipif(not os.path.exists("/root/reference/data/reference_double.h5"), reason="reference tree not mounted")
@pytest.mark.parametrize("precision", ["double", "single"])
def test_h5lite_reads_reference_files_like_fixtures(precision):
    f = File(f"/root/reference/data/reference_{precision}.h5")
    g = np.load(os.path.join(GOLDEN, f"reference_{precision}.npz"))
    assert sorted(f.keys()) == sorted(g.files)
    for name in g.files:
        assert np.array_equal(f[name], g[name]), name


@pytest.mark.skipif(not os.path.exists(INPUT_H5), reason="data/input.h5 is not shipped with the reference")
def test_oracle_reproduces_golden_when_input_is_available():
    """Point-wise pin of the oracle (config 1 of BASELINE.json): rel 1e-12 field-scaled."""
    from cloudsc2_b200 import iox

    op = iox.HDF5Operator(INPUT_H5)
    P = H.externals()
    for getter in ("get_yoethf_params", "get_yomcst_params", "get_yrecldp_params", "get_yrephli_params"):
        P.update(getattr(op, getter)().dict())
    f = op.f
    nz = op.get_nlev()

    def full(a):
        out = np.zeros((nz + 1, a.shape[1]))
        out[:nz] = a
        return out

    s = {
        "f_ap": full(f["PAP"]), "f_aph": np.array(f["PAPH"], dtype=np.float64), "f_lu": full(f["PLU"]),
        "f_lude": full(f["PLUDE"]), "f_mfd": full(f["PMFD"]), "f_mfu": full(f["PMFU"]), "f_q": full(f["PQ"]),
        "f_qi": full(f["PCLV"][1]), "f_ql": full(f["PCLV"][0]), "f_supsat": full(f["PSUPSAT"]), "f_t": full(f["PT"]),
        "f_tnd_cml_q": full(f["TENDENCY_CML_Q"]), "f_tnd_cml_t": full(f["TENDENCY_CML_T"]),
        "f_tnd_cml_qi": full(f["TENDENCY_CML_CLD"][1]), "f_tnd_cml_ql": full(f["TENDENCY_CML_CLD"][0]),
    }
    s = H.with_diagnostics(s, P)
    tn, dg = H.onp.cloudsc2_nl(s, op.get_timestep().total_seconds(), P)
    g = np.load(os.path.join(GOLDEN, "reference_double.npz"))
    pairs = {
        "PCLC": dg["f_clc"][:nz], "PFPLSN": dg["f_fplsn"], "PFHPSN": dg["f_fhpsn"], "PFPLSL": dg["f_fplsl"],
        "TENDENCY_LOC_T": tn["f_t"][:nz], "TENDENCY_LOC_Q": tn["f_q"][:nz],
    }
    for name, got in pairs.items():
        assert H.field_err(got, g[name]) < 1e-12 or not g[name].any(), name
