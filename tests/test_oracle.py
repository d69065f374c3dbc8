"""CPU: the NumPy oracle against the reference's own acceptance tests and the committed fixtures."""
import os

import numpy as np
import pytest

import helpers as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("block", ["base", "cold"])
def test_oracle_taylor_vshape(block):
    """Reference criterion (tangent_linear/validation.py:183-217): passed with penalty <= 5."""
    norms, _ = H.oracle_taylor(H.make_state(block), H.externals())
    passed, code, start = H.onp.taylor_score(norms)
    assert passed and code <= 5 and start <= 3, (norms, code, start)
    # V-shape: |1 - norm| shrinks ~10x per decade of factor2 on the way down, i.e. the remainder
    # NL(x + f2 dx) - NL(x) - f2 TL dx is O(f2^2) (slope 2), until round-off takes over
    err = np.abs(1 - norms)
    assert err[3] < 1e-3 and err.min() < 1e-6
    assert 0.05 < err[4] / err[3] < 0.2 and 0.05 < err[5] / err[4] < 0.2, err


@pytest.mark.parametrize("block,predicates", [("base", "tl"), ("cold", "tl"), ("cold", "reference")])
def test_oracle_symmetry(block, predicates):
    """Reference criterion (adjoint/validation.py:155-165): max_i norm3 < 1e4 eps."""
    _, _, n3, _ = H.oracle_symmetry(H.make_state(block), H.externals(LREGCL=True), predicates=predicates)
    assert n3.max() < 1e4, n3.max()


def test_oracle_symmetry_reference_predicates_break_on_rtt_crossing():
    """Documents WHY the product defaults to the TL predicates: the literal AD predicates are not
    the transpose of the TL where a level crosses RTT during the saturation adjustment."""
    _, _, n3, _ = H.oracle_symmetry(H.make_state("base"), H.externals(LREGCL=True), predicates="reference")
    assert n3.max() > 1e4


def test_oracle_ad_consumes_seeds():
    _, _, _, o = H.oracle_symmetry(H.make_state("cold", ncol=32), H.externals(LREGCL=True))
    for k in ("f_tnd_t_i", "f_tnd_q_i", "f_tnd_ql_i", "f_tnd_qi_i", "f_clc_i", "f_covptot_i", "f_fhpsl_i",
              "f_fhpsn_i", "f_fplsl_i", "f_fplsn_i"):
        assert not o["ad_in"][k].any(), k


@pytest.mark.parametrize("block", ["base", "cold"])
@pytest.mark.parametrize("precision,dtype", [("double", np.float64), ("single", np.float32)])
def test_oracle_matches_committed_fixtures(block, precision, dtype):
    """The oracle must not drift: fixtures were written by tests/golden/make_golden.py."""
    ref = np.load(os.path.join(GOLDEN, f"oracle_{block}_{precision}.npz"))
    got = H.oracle_run_all(block, dtype, 16)
    assert set(ref.files) == set(got)
    tol = 1e-13 if dtype == np.float64 else 1e-5  # libm differences between machines
    for name in ref.files:
        if name.startswith("sym_norm3"):
            continue
        r, g = ref[name], got[name]
        if np.abs(r).max() == 0:
            assert np.abs(g).max() == 0, name
        else:
            assert H.field_err(g, r) <= tol, (name, H.field_err(g, r))


def test_nl_invariants_like_golden():
    """The invariants the reference's golden outputs obey hold for the oracle."""
    P = H.externals()
    s = H.with_diagnostics(H.make_state("cold"), P)
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    assert np.array_equal(dg["f_fhpsn"], -dg["f_fplsn"] * P["RLSTT"])
    assert np.array_equal(dg["f_fhpsl"], -dg["f_fplsl"] * P["RLVTT"])
    assert not dg["f_covptot"].any()
    assert not dg["f_fplsn"][0].any() and not dg["f_fplsl"][0].any()
    assert dg["f_clc"].min() >= 0 and dg["f_clc"].max() <= 1
    assert not dg["f_fplsl"].any()  # all-cold columns: no rain, like PFPLSL == 0 in the golden file


def test_evaporation_branch_makes_covptot():
    P = H.externals(LEVAPLS2=True)
    s = H.with_diagnostics(H.make_state("base"), P)
    _, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    assert dg["f_covptot"].any()


def test_synthetic_block_exercises_both_sides_of_every_predicate():
    """Branch census on the base block: the parity tests are only as strong as the
    branches the inputs reach."""
    P = H.externals()
    s = H.with_diagnostics(H.make_state("base"), P)
    nz = 137
    dt = H.DT
    t = s["f_t"][:nz] + dt * s["f_tnd_cml_t"][:nz]
    tn, dg = H.onp.cloudsc2_nl(s, dt, P)
    clc = dg["f_clc"][:nz]
    both = lambda m: bool(m.any() and (~m).any())  # noqa: E731
    assert both(t < P["RTT"])                                   # p2 / p14 / p19 phase
    assert both(t < P["RTICE"])                                 # p6 ice supersaturation
    assert both(t > P["RTT"] + 2.0)                             # p12 melting possible
    assert (clc == 0).any() and (clc == 1).any() and ((clc > 0) & (clc < 1)).any()   # p7 three-way
    assert both(clc > P["ZEPS2"])                               # p13 autoconversion
    gdp = P["RG"] / (s["f_aph"][1:] - s["f_aph"][:-1])
    lude = dt * s["f_lude"][:nz] * gdp
    lo1 = (lude >= P["RLMIN"]) & (s["f_lu"][1:] >= P["ZEPS2"])
    assert both(lo1)                                            # p8 convective detrainment
    assert (dg["f_fplsl"] > 0).any() and (dg["f_fplsn"] > 0).any()   # rain and snow both occur
    # melting actually happens: rain flux appears below snow where t > RTT + 2
    snow_in, warm = dg["f_fplsn"][:nz] > 0, t > P["RTT"] + 2.0
    assert (snow_in & warm).any()
    # the tropopause rule fires for some columns and not for others
    trp = H.onp._trpaus(t, s["f_eta"], nz)
    assert both(trp > 0.1)
    # the saturation clip (esdp > ZQMAX) is reached at the model top
    foeew = P["R2ES"] * np.exp(P["R3IES"] * (t - P["RTT"]) / (t - P["R4IES"]))
    assert both(foeew / s["f_ap"][:nz] > P["ZQMAX"])
