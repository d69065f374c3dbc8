"""CPU: pins the oracle to OUTPUTS OF THE REFERENCE'S OWN SOURCE.

Three layers:
  1. the gtscript interpreter (oracle/gtscript_exec.py) on toy stencils whose results are known in closed form
     (forward / backward carries, k-offset reads of temporaries across computations, masked stores, inlined
     functions, compile-time branches, fp32 stays fp32) -- runs everywhere;
  2. `oracle/cloudsc2_numpy.py` (the hand restatement) against the reference's stencil files executed in place by
     that interpreter (oracle/ref_run.py), every stencil, fp64 and fp32, default and non-default flags -- runs where
     `/root/reference/src` exists (the authoring container; not the GPU box);
  3. the oracle against the committed outputs of (2), `tests/golden/ref_*.npz` -- runs everywhere, so the pin
     travels with the repository.

Tolerance: fp64 1e-13 field-scaled (measured: bit-identical on every field of every case); fp32 1e-5 (measured
<= 4e-6: NumPy's float32 scalar `**` and array `np.power` differ by an ulp in `scalm`)."""
import os

import numpy as np
import pytest

import helpers as H
from oracle import gtscript_exec as gx
from oracle import ref_run

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_reference = pytest.mark.skipif(not ref_run.available(), reason="/root/reference/src not present on this box")
TOL = {np.dtype(np.float64): 1e-13, np.dtype(np.float32): 1e-5}


# ------------------------------------------------------------------------------------------------------------
# 1. interpreter semantics on toy stencils (names below are what gtscript injects; never called as python)
# ------------------------------------------------------------------------------------------------------------
class gtscript:  # annotation source only
    class _F:
        def __getitem__(self, item):
            return self

    Field = _F()
    K = IJ = None


def toy_sweeps(in_a: gtscript.Field["float"], out_fwd: gtscript.Field["float"], out_bwd: gtscript.Field["float"],
               out_shift: gtscript.Field["float"], tmp_c: gtscript.Field[gtscript.IJ, "float"], *, w: "float"):
    from __externals__ import SCALE

    with computation(FORWARD), interval(0, 1):
        tmp_c[0, 0] = 0.0
    with computation(FORWARD), interval(0, -1):
        tmp_c[0, 0] = tmp_c[0, 0] + w * in_a[0, 0, 0]
        out_fwd[0, 0, 0] = SCALE * tmp_c[0, 0]
        loc = in_a[0, 0, 0] * 2.0
    with computation(BACKWARD):
        with interval(-1, None):
            out_bwd[0, 0, 0] = 0.0
        with interval(0, -1):
            out_bwd[0, 0, 0] = out_bwd[0, 0, 1] + in_a[0, 0, 0]
    with computation(FORWARD):
        with interval(0, 1):
            out_shift[0, 0, 0] = -1.0
        with interval(1, None):
            out_shift[0, 0, 0] = loc[0, 0, -1]


def test_interpreter_sweeps_carries_and_temporaries():
    nx, nk = 5, 7
    rng = np.random.default_rng(0)
    a = rng.random((nk, nx))
    fwd, bwd, shift, c = np.zeros((nk, nx)), np.zeros((nk, nx)), np.zeros((nk, nx)), np.zeros(nx)
    st = gx.Stencil("toy_sweeps", {"SCALE": 3.0}, np.float64, fn=toy_sweeps)
    st(in_a=a, out_fwd=fwd, out_bwd=bwd, out_shift=shift, tmp_c=c, w=0.5, origin=(0, 0, 0), domain=(nx, 1, nk))
    np.testing.assert_allclose(fwd[: nk - 1], 3.0 * 0.5 * np.cumsum(a[: nk - 1], axis=0), rtol=1e-15)
    assert not fwd[nk - 1].any()  # interval(0, -1) leaves the last level alone
    np.testing.assert_allclose(bwd[: nk - 1], np.cumsum(a[nk - 2 :: -1], axis=0)[::-1], rtol=1e-15)
    assert np.array_equal(shift[0], np.full(nx, -1.0))
    assert np.array_equal(shift[1:], 2.0 * a[:-1])  # 3-D temporary read one level up, in a later computation
    np.testing.assert_allclose(c, 0.5 * a[: nk - 1].sum(axis=0), rtol=1e-15)  # IJ field persists across levels


def toy_fn_inner(x, y):
    from __externals__ import FLAG

    if FLAG == 0:
        z = x + y
        return z, x - y


def toy_fn(a, b):
    if a > b:
        hi = a
        lo = b
    else:
        hi = b
        lo = a
    s, d = toy_fn_inner(hi, lo)
    s, d = toy_fn_inner(s, d)
    return s, d


def toy_masks(in_a: gtscript.Field["float"], in_b: gtscript.Field["float"], in_lev: gtscript.Field[gtscript.K, "float"],
              out_x: gtscript.Field["float"], out_y: gtscript.Field["float"], out_z: gtscript.Field["float"]):
    from __externals__ import FLAG, THRESH

    with computation(PARALLEL), interval(...):
        if FLAG == 1:
            out_x[0, 0, 0] = -99.0
        else:
            out_x[0, 0, 0] = 1.0
        if in_a[0, 0, 0] > THRESH:
            out_x[0, 0, 0] += 10.0
            if in_b[0, 0, 0] > THRESH and in_lev[0] < 2.5:
                only_here = 5.0
        elif in_a[0, 0, 0] > 0.25:
            out_x[0, 0, 0] -= 10.0
        out_y[0, 0, 0] = only_here  # zero where never assigned (zero-initialised temporary)
        s, d = toy_fn(in_a, in_b)
        out_z[0, 0, 0] = s * 1000.0 + d


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_interpreter_masks_functions_and_dtype(dtype):
    nx, nk = 11, 4
    rng = np.random.default_rng(1)
    a, b = rng.random((nk, nx)).astype(dtype), rng.random((nk, nx)).astype(dtype)
    lev = np.arange(nk, dtype=dtype)
    x, y, z = (np.full((nk, nx), 7.0, dtype=dtype) for _ in range(3))
    st = gx.Stencil("toy_masks", {"FLAG": 0, "THRESH": 0.5}, dtype, fn=toy_masks)
    # functions resolve through the globals of the module that defines the stencil
    st(in_a=a, in_b=b, in_lev=lev, out_x=x, out_y=y, out_z=z, origin=(0, 0, 0), domain=(nx, 1, nk))
    assert x.dtype == dtype and z.dtype == dtype
    exp_x = np.where(a > 0.5, 11.0, np.where(a > 0.25, -9.0, 1.0))
    assert np.array_equal(x, exp_x.astype(dtype))
    exp_y = np.where((a > 0.5) & (b > 0.5) & (lev[:, None] < 2.5), 5.0, 0.0)
    assert np.array_equal(y, exp_y.astype(dtype))
    hi, lo = np.maximum(a, b), np.minimum(a, b)
    s1, d1 = hi + lo, hi - lo
    s2, d2 = s1 + d1, s1 - d1
    assert np.array_equal(z, s2 * dtype(1000.0) + d2)


def toy_promotes(in_a: gtscript.Field["float"], out_x: gtscript.Field["float"], in_wide: gtscript.Field[gtscript.K, "float"]):
    with computation(PARALLEL), interval(...):
        out_x[0, 0, 0] = in_a[0, 0, 0] * in_wide[0]


def test_interpreter_refuses_silent_promotion():
    a = np.ones((3, 2), dtype=np.float32)
    st = gx.Stencil("toy_promotes", {}, np.float32, fn=toy_promotes)
    with pytest.raises(TypeError):
        st(in_a=a, out_x=np.zeros_like(a), in_wide=np.ones(3, dtype=np.float64), origin=(0, 0, 0), domain=(2, 1, 3))


def toy_parallel_offset(in_a: gtscript.Field["float"], out_x: gtscript.Field["float"]):
    with computation(PARALLEL), interval(0, -1):
        t = in_a[0, 0, 0]
        out_x[0, 0, 0] = t[0, 0, 1]


def test_interpreter_rejects_parallel_blocks_it_cannot_run_level_by_level():
    a = np.ones((3, 2))
    st = gx.Stencil("toy_parallel_offset", {}, np.float64, fn=toy_parallel_offset)
    with pytest.raises(AssertionError):
        st(in_a=a, out_x=np.zeros_like(a), origin=(0, 0, 0), domain=(2, 1, 3))


# ------------------------------------------------------------------------------------------------------------
# 2. the oracle against the reference's stencil sources executed here
# ------------------------------------------------------------------------------------------------------------
@needs_reference
def test_reference_files_are_loaded_unmodified_from_the_reference_tree():
    gx.load_reference()
    assert set(gx.STENCILS) >= {"saturation", "cloudsc2_nl", "cloudsc2_tl", "cloudsc2_ad", "state_increment",
                                "perturbed_state"}
    assert set(gx.FUNCTIONS) >= {"f_foealfa", "f_foeewm", "f_foeewmcu", "f_cuadjtqs_nl", "f_cuadjtqs_tl", "f_cuadjtqs_ad"}
    for name in ("cloudsc2_nl", "cloudsc2_tl", "cloudsc2_ad", "saturation"):
        assert gx.source_file(name).startswith(gx.REFERENCE_SRC), gx.source_file(name)
    import sys

    assert "gt4py" not in sys.modules and "ifs_physics_common" not in sys.modules  # stubs do not linger


@needs_reference
def test_component_glue_is_read_from_the_reference_components():
    """Externals and the state-name -> stencil-argument map come from the reference's component classes."""
    P = H.externals(LREGCL=False, LEVAPLS2=True)
    name, ext = ref_run.component_externals("Cloudsc2TL", P, dict(lphylin=True, ldrain1d=False), 137)
    assert name == "cloudsc2_tl" and ext["NLEV"] == 137 and ext["LREGCL"] is False and ext["LEVAPLS2"] is True
    assert ext["ZQMAX"] == 0.5 and ext["ICALL"] == 0 and ext["RTT"] == P["RTT"]
    name, ext = ref_run.component_externals("Saturation", P, dict(kflag=0, lphylin=False), 137)
    assert name == "saturation" and ext["KFLAG"] == 0 and ext["QMAX"] == 0.5 and ext["LPHYLIN"] is False
    kw, half = ref_run.stencil_call_map("Cloudsc2AD")
    assert half and kw["in_tnd_t_i"] == ("state", "f_tnd_t_i") and kw["out_tnd_cml_q_i"] == ("out_tendencies", "f_cml_q_i")
    assert kw["tmp_klevel"] == ("klevel", None) and kw["dt"] == ("dt", None) and kw["tmp_rfln_i"] == ("tmp_ij", None)
    kw, half = ref_run.stencil_call_map("Saturation")
    assert not half  # common/saturation.py:73: full levels only


def _compare(got, ref, tol, what):
    exact = True
    for k, r in ref.items():
        g = got[k]
        assert g.shape == r.shape and g.dtype == r.dtype, (what, k)
        if not np.array_equal(g, r):
            exact = False
            if np.max(np.abs(r)) == 0:
                assert np.max(np.abs(g)) == 0, (what, k)
            else:
                assert H.field_err(g, r) <= tol, (what, k, H.field_err(g, r))
    return exact


CASES = [(name,) + spec for name, spec in H.REF_FIXTURES.items()]


@needs_reference
@pytest.mark.parametrize("name,block,dtype,ncol,flags", CASES, ids=[c[0] for c in CASES])
def test_oracle_equals_reference_source(name, block, dtype, ncol, flags):
    """saturation, NL, state_increment, perturbed_state, TL, AD (+ consumed seeds) -- 100 fields per case."""
    ncol = 100 if not flags else ncol  # the two default-flag blocks in full
    ref = H.pipeline_run_all(ref_run, block, dtype, ncol, **flags)
    got = H.pipeline_run_all(H.onp, block, dtype, ncol, ad_kwargs=dict(predicates="reference"), **flags)
    assert set(ref) == set(got)
    exact = _compare(got, ref, TOL[np.dtype(dtype)], name)
    if np.dtype(dtype) == np.float64:
        assert exact, f"{name}: the fp64 oracle is no longer bit-identical to the reference's source"
    for k in H.SEED_KEYS:
        assert not ref["ad_seed_" + k].any(), f"reference AD left seed {k} non-zero"


@needs_reference
def test_reference_source_passes_its_own_taylor_test():
    """tangent_linear/validation.py:150-217 with every stencil call going to the reference's source."""
    P = H.externals(LREGCL=False)
    st = {k: np.ascontiguousarray(v[:, :40]) for k, v in H.make_state("base").items()}
    s = dict(st, f_eta=H.onp.eta_levels(st["f_ap"], st["f_aph"]))
    s["f_qsat"] = ref_run.saturation(s["f_ap"], s["f_t"], P)
    tn, dg = ref_run.cloudsc2_nl(s, H.DT, P)
    s.update(ref_run.state_increment(s, 0.01))
    ttl, dtl = ref_run.cloudsc2_tl(s, H.DT, P)
    norms = []
    for i in range(10):
        f2 = float(10 ** -(i + 1))
        sp = dict(ref_run.perturbed_state(s, f2), f_eta=s["f_eta"])
        tnp_, dgp = ref_run.cloudsc2_nl(sp, H.DT, P)
        norms.append(H.onp.taylor_norm(f2, tn, dg, tnp_, dgp, ttl, dtl))
    passed, code, start = H.onp.taylor_score(np.array(norms))
    assert passed and code <= 5 and start <= 3, (norms, code, start)


# ------------------------------------------------------------------------------------------------------------
# 3. the oracle against the committed outputs of the reference's source (travels to every box)
# ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,block,dtype,ncol,flags", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_committed_reference_outputs(name, block, dtype, ncol, flags):
    ref = np.load(os.path.join(GOLDEN, name + ".npz"))
    got = H.pipeline_run_all(H.onp, block, dtype, ncol, ad_kwargs=dict(predicates="reference"), **flags)
    assert set(ref.files) == set(got)
    tol = TOL[np.dtype(dtype)]
    if flags.get("LEVAPLS2") or flags.get("LDRAIN1D"):
        tol = 1e-10  # the reference's AD of the evaporation branch grows adjoints to 1e80: libm ulps get amplified
    _compare(got, {k: ref[k] for k in ref.files}, tol, name)
