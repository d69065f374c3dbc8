"""CPU: the product's column code (csrc/cs2_columns.cuh, compiled for the host by oracle/Makefile)
against the independent NumPy oracle.  This checks the kernel MATH without a GPU; the `-m gpu`
tests check the same thing through the CUDA kernels."""
import numpy as np
import pytest

import helpers as H

DTYPES = [np.float64, np.float32]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("block", ["base", "cold"])
def test_twin_saturation_and_nl(block, dtype):
    P = H.externals()
    s = H.with_diagnostics(H.make_state(block, dtype), P)
    tol = H.TOL[np.dtype(dtype)]
    assert H.field_err(H.twin_saturation(s["f_ap"], s["f_t"], P), s["f_qsat"]) <= tol
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    ttn, tdg = H.twin_nl(s, H.DT, P)
    H.assert_fields_close(ttn, tn, tol, "NL tendencies: ")
    H.assert_fields_close(tdg, dg, tol, "NL diagnostics: ")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("block", ["base", "cold"])
@pytest.mark.parametrize("flags", [dict(), dict(LPHYLIN=False), dict(RVTMP2=0.61)])
def test_twin_nl_split(block, dtype, flags):
    """The two half-level functions of the split NL kernel (A: carry-independent, B: carry-dependent) give NL."""
    P = H.externals(**flags)
    s = H.with_diagnostics(H.make_state(block, dtype), P)
    tol = H.TOL[np.dtype(dtype)]
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    ttn, tdg = H.twin_nl(s, H.DT, P, split=True)
    H.assert_fields_close(ttn, tn, tol, f"NL split {flags} tendencies: ")
    H.assert_fields_close(tdg, dg, tol, f"NL split {flags} diagnostics: ")


@pytest.mark.parametrize("flags", [dict(LEVAPLS2=True), dict(LDRAIN1D=True), dict(LPHYLIN=False),
                                    dict(LPHYLIN=False, LEVAPLS2=True)])
def test_twin_nl_flag_paths(flags):
    """Evaporation branch and the non-LPHYLIN thermodynamics (NL only)."""
    P = H.externals(**flags)
    s = H.with_diagnostics(H.make_state("base"), P)
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    ttn, tdg = H.twin_nl(s, H.DT, P)
    H.assert_fields_close(ttn, tn, 1e-12, f"NL {flags}: ")
    H.assert_fields_close(tdg, dg, 1e-12, f"NL {flags}: ")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("block", ["base", "cold"])
@pytest.mark.parametrize("lregcl", [True, False])
def test_twin_tl(block, dtype, lregcl):
    P = H.externals(LREGCL=lregcl)
    s = H.with_diagnostics(H.make_state(block, dtype), P)
    s.update(H.onp.state_increment(s, 0.01))
    rt, rd = H.onp.cloudsc2_tl(s, H.DT, P)
    tt, td = H.twin_tl(s, H.DT, P)
    tol_t = tol_d = H.TOL[np.dtype(dtype)]
    if dtype == np.float32:
        s64 = {k: v.astype(np.float64) for k, v in s.items()}
        rt64, rd64 = H.onp.cloudsc2_tl(s64, H.DT, P)
        tol_t, tol_d = H.fp32_field_tolerances(rt, rt64), H.fp32_field_tolerances(rd, rd64)
    H.assert_fields_close(tt, rt, tol_t, "TL tendencies: ")
    H.assert_fields_close(td, rd, tol_d, "TL diagnostics: ")


@pytest.mark.parametrize("flags", [dict(LEVAPLS2=True), dict(LDRAIN1D=True)])
@pytest.mark.parametrize("lregcl", [True, False])
@pytest.mark.parametrize("block", ["base", "cold"])
def test_twin_tl_evaporation_branch(block, lregcl, flags):
    """TL with the precipitation-evaporation branch (LEVAPLS2 / LDRAIN1D) against the oracle's literal restatement of
    tangent_linear/_stencils/cloudsc2.py:525-616 (the branch is exercised: f_covptot and f_covptot_i are non-zero)."""
    P = H.externals(LREGCL=lregcl, **flags)
    s = H.with_diagnostics(H.make_state(block), P)
    s.update(H.onp.state_increment(s, 0.01))
    rt, rd = H.onp.cloudsc2_tl(s, H.DT, P)
    assert np.count_nonzero(rd["f_covptot"]) > 0
    if block == "base":
        assert np.count_nonzero(rd["f_covptot_i"]) > 0
    tt, td = H.twin_tl(s, H.DT, P)
    H.assert_close_except_total_evaporation_knife_edges({**tt, **td}, {**rt, **rd}, 1e-12, max_columns=2, what=f"TL {flags}: ")


@pytest.mark.parametrize("flags", [dict(LEVAPLS2=True), dict(LDRAIN1D=True)])
@pytest.mark.parametrize("predicates", ["tl", "reference"])
@pytest.mark.parametrize("block", ["base", "cold"])
def test_twin_ad_evaporation_branch(block, predicates, flags):
    """AD with the precipitation-evaporation branch against the oracle's literal restatement of
    adjoint/_stencils/cloudsc2.py:635-719,808-817,936-941, seeded with the TL outputs like the symmetry harness."""
    P = H.externals(LREGCL=True, **flags)
    _, _, _, o = H.oracle_symmetry(H.make_state(block), P, predicates=predicates)
    assert np.count_nonzero(o["diags_tl"]["f_covptot_i"]) > 0 or block == "cold"
    ad_in = dict(o["state"])
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = o["tends_tl"][f"f_{x}_i"].copy()
    for k, v in o["diags_tl"].items():
        ad_in[k] = v.copy()
    tad, dad, consumed = H.twin_ad(ad_in, H.DT, P, predicates=predicates)
    H.assert_close_except_total_evaporation_knife_edges({**tad, **dad}, {**o["tends_ad"], **o["diags_ad"]}, 1e-12,
                                                        max_columns=2, what=f"AD {flags} {predicates}: ")
    for k, v in consumed.items():
        assert not v.any(), f"seed {k} not zeroed"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("block,predicates", [("base", "tl"), ("base", "reference"), ("cold", "tl")])
def test_twin_ad(block, predicates, dtype):
    P = H.externals(LREGCL=True)
    _, _, _, o = H.oracle_symmetry(H.make_state(block, dtype), P, predicates=predicates)
    ad_in = dict(o["state"])
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = o["tends_tl"][f"f_{x}_i"].copy()
    for k, v in o["diags_tl"].items():
        ad_in[k] = v.copy()
    tad, dad, consumed = H.twin_ad(ad_in, H.DT, P, predicates=predicates)
    tol_t = tol_d = H.TOL[np.dtype(dtype)]
    if dtype == np.float32:
        st64 = {k: v.astype(np.float64) for k, v in H.make_state(block, dtype).items()}
        _, _, _, o64 = H.oracle_symmetry(st64, P, predicates=predicates)
        tol_t = H.fp32_field_tolerances(o["tends_ad"], o64["tends_ad"])
        tol_d = H.fp32_field_tolerances(o["diags_ad"], o64["diags_ad"])
    H.assert_fields_close(tad, o["tends_ad"], tol_t, "AD tendencies: ")
    H.assert_fields_close(dad, o["diags_ad"], tol_d, "AD diagnostics: ")
    for k, v in consumed.items():
        assert not v.any(), f"seed {k} not zeroed"


def test_twin_symmetry_at_roundoff():
    """<TL x, TL x> = <x, AD TL x> per column, with the product's own TL and AD."""
    P = H.externals(LREGCL=True)
    s = H.with_diagnostics(H.make_state("base"), P)
    si = H.onp.state_increment(s, 0.01, ignore_supsat=True)
    s.update(si)
    tt, td = H.twin_tl(s, H.DT, P)
    n1 = H.onp.symmetry_norm1(tt, td)
    ad_in = dict(s)
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = tt[f"f_{x}_i"].copy()
    ad_in.update({k: v.copy() for k, v in td.items()})
    tad, dad, _ = H.twin_ad(ad_in, H.DT, P, predicates="tl")
    n3 = H.onp.symmetry_norm3(n1, H.onp.symmetry_norm2(si, tad, dad), np.float64)
    assert n3.max() < 1e4, n3.max()


def test_twin_tiled_blocks_are_bit_identical():
    P = H.externals()
    s = H.with_diagnostics(H.make_state("base", ncol=300), P)
    tn, dg = H.twin_nl(s, H.DT, P)
    for d in (tn, dg):
        for k, v in d.items():
            assert np.array_equal(v[:, :100], v[:, 100:200]) and np.array_equal(v[:, :100], v[:, 200:300]), k


@pytest.mark.parametrize("ncol", [1, 31, 33])
def test_twin_ragged_column_counts(ncol):
    P = H.externals()
    s = H.with_diagnostics(H.make_state("base", ncol=ncol), P)
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    ttn, tdg = H.twin_nl(s, H.DT, P)
    H.assert_fields_close(ttn, tn, 1e-12)
    H.assert_fields_close(tdg, dg, 1e-12)


@pytest.mark.parametrize("seed", [3, 11, 29])
def test_twin_random_blocks_nl_tl_ad(seed):
    """Other seeds of the synthetic generator (different branch patterns): NL, TL and AD of the kernel code vs oracle."""
    P = H.externals(LREGCL=True)
    st = H.make_state("base", np.float64, 100, seed=seed)
    _, _, n3, o = H.oracle_symmetry(st, P, predicates="tl")
    s = o["state"]
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    ttn, tdg = H.twin_nl(s, H.DT, P)
    H.assert_fields_close(ttn, tn, 1e-12, f"seed {seed} NL: ")
    H.assert_fields_close(tdg, dg, 1e-12, f"seed {seed} NL: ")
    tt, td = H.twin_tl(s, H.DT, P)
    H.assert_fields_close(tt, o["tends_tl"], 1e-12, f"seed {seed} TL: ")
    H.assert_fields_close(td, o["diags_tl"], 1e-12, f"seed {seed} TL: ")
    ad_in = dict(s)
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = o["tends_tl"][f"f_{x}_i"].copy()
    ad_in.update({k: v.copy() for k, v in o["diags_tl"].items()})
    tad, dad, _ = H.twin_ad(ad_in, H.DT, P, predicates="tl")
    H.assert_fields_close(tad, o["tends_ad"], 1e-12, f"seed {seed} AD: ")
    H.assert_fields_close(dad, o["diags_ad"], 1e-12, f"seed {seed} AD: ")
    assert n3.max() < 1e4


@pytest.mark.parametrize("nz,dt", [(60, 900.0), (137, 1200.0), (20, 3600.0)])
def test_twin_other_level_counts_and_timesteps(nz, dt):
    """Nothing in the kernels is specialised to 137 levels or dt = 3600 s."""
    from cloudsc2_b200 import synthetic

    P = H.externals(LREGCL=True)
    st = {k: np.ascontiguousarray(v) for k, v in synthetic.base_block(nz=nz, ncol=64, seed=5).items()}
    s = H.with_diagnostics(st, P)
    tn, dg = H.onp.cloudsc2_nl(s, dt, P)
    ttn, tdg = H.twin_nl(s, dt, P)
    H.assert_fields_close(ttn, tn, 1e-12, f"nz={nz} NL: ")
    H.assert_fields_close(tdg, dg, 1e-12, f"nz={nz} NL: ")
    s.update(H.onp.state_increment(s, 0.01, ignore_supsat=True))
    rt, rd = H.onp.cloudsc2_tl(s, dt, P)
    tt, td = H.twin_tl(s, dt, P)
    H.assert_fields_close(tt, rt, 1e-12, f"nz={nz} TL: ")
    H.assert_fields_close(td, rd, 1e-12, f"nz={nz} TL: ")
    ad_in = dict(s)
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = rt[f"f_{x}_i"].copy()
    ad_in.update({k: v.copy() for k, v in rd.items()})
    ref_in = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in ad_in.items()}
    rat, rad = H.onp.cloudsc2_ad(ref_in, dt, P, predicates="tl")
    tad, dad, _ = H.twin_ad(ad_in, dt, P, predicates="tl")
    H.assert_fields_close(tad, rat, 1e-12, f"nz={nz} AD: ")
    H.assert_fields_close(dad, rad, 1e-12, f"nz={nz} AD: ")


@pytest.mark.parametrize("flags", [dict(LPHYLIN=True, KFLAG=1), dict(LPHYLIN=False, KFLAG=1), dict(LPHYLIN=False, KFLAG=0)])
@pytest.mark.parametrize("dtype", DTYPES)
def test_twin_saturation_flag_paths(flags, dtype):
    """All three branches of the saturation stencil (common/_stencils/saturation.py:30-41)."""
    P = H.externals(**flags)
    st = H.make_state("base", dtype)
    ref = H.onp.saturation(st["f_ap"], st["f_t"], P)
    got = H.twin_saturation(st["f_ap"], st["f_t"], P)
    assert H.field_err(got, ref) <= H.TOL[np.dtype(dtype)]
    assert not got[137].any()  # the padding level is outside the stencil's domain


@pytest.mark.parametrize("nz", [1, 2, 3, 7])
def test_twin_tiny_level_counts(nz):
    """Degenerate columns (1-7 levels: no tropopause window, the level below the bottom level is the padding level):
    NL, TL and AD still equal the oracle."""
    from cloudsc2_b200 import synthetic

    P = H.externals(LREGCL=True)
    st = {k: np.ascontiguousarray(v) for k, v in synthetic.base_block(nz=nz).items()}
    _, _, _, ref = H.oracle_symmetry(st, P, predicates="tl")
    s = ref["state"]
    tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    ttn, tdg = H.twin_nl(s, H.DT, P)
    H.assert_fields_close(ttn, tn, 1e-12, "NL: ")
    H.assert_fields_close(tdg, dg, 1e-12, "NL: ")
    tt, td = H.twin_tl(s, H.DT, P)
    H.assert_fields_close(tt, ref["tends_tl"], 1e-12, "TL: ")
    H.assert_fields_close(td, ref["diags_tl"], 1e-12, "TL: ")
    ad_in = dict(s)
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = ref["tends_tl"][f"f_{x}_i"].copy()
    for k, v in ref["diags_tl"].items():
        ad_in[k] = v.copy()
    tad, dad, _ = H.twin_ad(ad_in, H.DT, P, predicates="tl")
    H.assert_fields_close(tad, ref["tends_ad"], 1e-12, "AD: ")
    H.assert_fields_close(dad, ref["diags_ad"], 1e-12, "AD: ")
