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
