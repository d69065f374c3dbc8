"""Shared test helpers: oracle harness, host-twin runner, parity metric."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Dict

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
PKG_DIR = os.path.join(ROOT, "gt4py-dwarf-p-cloudsc2-tl-ad_b200")
for _p in (ROOT, PKG_DIR):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from cloudsc2_b200 import _lib, iox, synthetic  # noqa: E402
import oracle.cloudsc2_numpy as onp  # noqa: E402

DT = 3600.0
TWIN_PATH = os.path.join(ROOT, "oracle", "_build", "libcs2_host_twin.so")

# Tolerances of BASELINE.json `north_star`: field-scaled max error
TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def externals(**overrides) -> Dict:
    d = iox.ifs_defaults()
    P: Dict = {}
    for v in d.values():
        P.update(v.dict())
    P.update(ICALL=0, LPHYLIN=True, LDRAIN1D=False, ZEPS1=1e-12, ZEPS2=1e-10, ZQMAX=0.5, ZSCAL=0.9, KFLAG=1, QMAX=0.5)
    P.update(overrides)
    return P


def make_state(block: str = "base", dtype=np.float64, ncol: int = 100, seed: int = 0) -> Dict[str, np.ndarray]:
    blk = synthetic.base_block(seed=seed) if block == "base" else synthetic.cold_block(seed=seed + 1)
    if ncol != synthetic.KLON:
        blk = synthetic.tile(blk, ncol)
    return {k: np.ascontiguousarray(v.astype(dtype)) for k, v in blk.items()}


def with_diagnostics(state: Dict[str, np.ndarray], P: Dict) -> Dict[str, np.ndarray]:
    s = dict(state)
    s["f_eta"] = onp.eta_levels(s["f_ap"], s["f_aph"])
    s["f_qsat"] = onp.saturation(s["f_ap"], s["f_t"], P)
    return s


def field_err(a: np.ndarray, b: np.ndarray) -> float:
    """max|a-b| / max|b| (field-scaled)."""
    scale = max(float(np.max(np.abs(b))), np.finfo(np.float64).tiny)
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)))) / scale


def fp32_field_tolerances(ref32: Dict[str, np.ndarray], ref64: Dict[str, np.ndarray], base: float = 1e-5,
                          factor: float = 4.0) -> Dict[str, float]:
    """Per-field tolerance for fp32 runs: BASELINE's 1e-5, widened -- only where needed -- to `factor` x the
    error the reference's OWN fp32 execution (the fp32 oracle) has against the fp64 oracle on the same
    inputs.  The TL/AD perturbation fields are differences of nearly equal quantities; there the fp32 oracle
    itself is only accurate to ~1e-5 of the field maximum, so two correct fp32 implementations (different
    libm, different but equivalent operation order) differ from each other by a small multiple of that.
    (Same for cloud cover near clc -> 0, where one fp32 ulp of qsat moves clc by > 1e-5.)"""
    tol = {}
    for name, r32 in ref32.items():
        r64 = ref64.get(name)
        own = field_err(r32, r64) if r64 is not None and np.max(np.abs(r64)) > 0 else 0.0
        tol[name] = max(base, factor * own)
    return tol


def assert_close_except_total_evaporation_knife_edges(got, ref, tol: float, max_columns: int, what: str = "") -> int:
    """assert_fields_close for runs with the precipitation-evaporation branch on; `got` / `ref` hold tendencies AND
    diagnostics.  Returns the number of columns left out.

    When ALL precipitation of a level evaporates, the reference computes `sfln - dpr * sfln / prtot` with
    dpr == prtot, which is an exact 0 or a +-1e-24 round-off residue depending on the last bits of sfln, and the
    level below tests `sfl != 0` (nonlinear/_stencils/cloudsc2.py:238) to decide whether that "snow" melts: the
    perturbation then leaves the level as rain or as snow.  Two correct implementations whose fluxes differ by an
    ulp can land on different sides.  A column may therefore miss the tolerance only if it shows such a residue in
    one of the two runs, and at most `max_columns` columns may."""
    tot_r = ref["f_fplsl"] + ref["f_fplsn"]
    tot_g = got["f_fplsl"] + got["f_fplsn"]
    residue = ((tot_g != 0) & (np.abs(tot_g) < 1e-18)) | ((tot_r != 0) & (np.abs(tot_r) < 1e-18)) | ((tot_g == 0) != (tot_r == 0))
    flagged = np.unique(np.nonzero(residue)[1])
    failing = set()
    for name, r in ref.items():
        g = got[name]
        assert g.shape == r.shape and np.all(np.isfinite(g)), f"{what}{name}"
        scale = np.max(np.abs(r))
        if scale == 0.0:
            cols = np.nonzero(np.max(np.abs(g), axis=0) != 0.0)[0]
        else:
            cols = np.nonzero(np.max(np.abs(g - r), axis=0) / scale > tol)[0]
        failing.update(int(c) for c in cols)
    stray = sorted(failing - set(int(c) for c in flagged))
    assert not stray, f"{what}columns {stray[:10]} miss the tolerance {tol:g} without a total-evaporation residue"
    assert len(failing) <= max_columns, f"{what}{len(failing)} knife-edge columns: {sorted(failing)[:10]}"
    return len(failing)


def assert_fields_close(got: Dict[str, np.ndarray], ref: Dict[str, np.ndarray], tol, what: str = "") -> None:
    """`tol`: one number, or a dict of per-field tolerances (see fp32_field_tolerances)."""
    bad = []
    tols = tol if isinstance(tol, dict) else None
    for name, r in ref.items():
        if tols is not None:
            tol = tols[name]
        g = got[name]
        assert g.shape == r.shape, f"{what}{name}: shape {g.shape} != {r.shape}"
        if not np.all(np.isfinite(g)):
            bad.append(f"{name}: non-finite values")
            continue
        if np.max(np.abs(r)) == 0.0:
            if np.max(np.abs(g)) != 0.0:
                bad.append(f"{name}: reference is identically 0, got max {np.max(np.abs(g)):.3e}")
            continue
        e = field_err(g, r)
        if e > tol:
            bad.append(f"{name}: field-scaled error {e:.3e} > {tol:.1e}")
    assert not bad, what + "; ".join(bad)


# ------------------------------------------------------------------------------------------
# oracle pipelines (the reference's drivers restated on the oracle)
# ------------------------------------------------------------------------------------------
def oracle_taylor(state, P, dt=DT, f1=0.01, nf2=10):
    """tangent_linear/validation.py:150-181 on the oracle."""
    P = dict(P, LREGCL=False)
    s = with_diagnostics(state, P)
    tn, dg = onp.cloudsc2_nl(s, dt, P)
    s.update(onp.state_increment(s, f1))
    ttl, dtl = onp.cloudsc2_tl(s, dt, P)
    norms = []
    for i in range(nf2):
        f2 = float(10 ** -(i + 1))
        sp = onp.perturbed_state(s, f2)
        sp["f_eta"] = s["f_eta"]
        tnp_, dgp = onp.cloudsc2_nl(sp, dt, P)
        norms.append(onp.taylor_norm(f2, tn, dg, tnp_, dgp, ttl, dtl))
    return np.array(norms), dict(tends_nl=tn, diags_nl=dg, tends_tl=ttl, diags_tl=dtl)


def oracle_symmetry(state, P, dt=DT, f=0.01, predicates="tl"):
    """adjoint/validation.py:132-165 on the oracle."""
    s = with_diagnostics(state, P)
    si = onp.state_increment(s, f, ignore_supsat=True)
    s.update(si)
    ttl, dtl = onp.cloudsc2_tl(s, dt, P)
    n1 = onp.symmetry_norm1(ttl, dtl)
    ad_in = dict(s)
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = ttl[f"f_{x}_i"].copy()
    for k, v in dtl.items():
        ad_in[k] = v.copy()
    tad, dad = onp.cloudsc2_ad(ad_in, dt, P, predicates=predicates)
    n2 = onp.symmetry_norm2(si, tad, dad)
    n3 = onp.symmetry_norm3(n1, n2, s["f_ap"].dtype)
    return n1, n2, n3, dict(state=s, tends_tl=ttl, diags_tl=dtl, tends_ad=tad, diags_ad=dad, ad_in=ad_in)


def oracle_run_all(block: str, dtype, ncol: int) -> Dict[str, np.ndarray]:
    """NL, TL and AD oracle outputs on the first `ncol` columns of a synthetic block, flattened
    into one dict for the golden fixtures (tests/golden/make_golden.py)."""
    P = externals(LREGCL=True)
    st = {k: np.ascontiguousarray(v[:, :ncol]) for k, v in make_state(block, dtype).items()}
    # eta must come from the block's column 0, which is kept
    s = with_diagnostics(st, P)
    out = {"in_" + k: v for k, v in s.items()}
    tn, dg = onp.cloudsc2_nl(s, DT, P)
    out.update({"nl_t_" + k: v for k, v in tn.items()})
    out.update({"nl_d_" + k: v for k, v in dg.items()})
    for pred in ("reference", "tl"):
        _, _, n3, o = oracle_symmetry(st, P, predicates=pred)
        if pred == "tl":
            out.update({"tl_t_" + k: v for k, v in o["tends_tl"].items()})
            out.update({"tl_d_" + k: v for k, v in o["diags_tl"].items()})
        out.update({f"ad_{pred}_t_" + k: v for k, v in o["tends_ad"].items()})
        out.update({f"ad_{pred}_d_" + k: v for k, v in o["diags_ad"].items()})
        out[f"sym_norm3_{pred}"] = n3
    return out


SEED_KEYS = ("f_tnd_t_i", "f_tnd_q_i", "f_tnd_ql_i", "f_tnd_qi_i", "f_clc_i", "f_covptot_i", "f_fhpsl_i", "f_fhpsn_i",
             "f_fplsl_i", "f_fplsn_i")  # adjoint/microphysics.py:106-120


def pipeline_run_all(impl, block: str, dtype, ncol: int, ad_kwargs=None, **flags) -> Dict[str, np.ndarray]:
    """saturation -> NL -> state_increment(0.01, ignore_supsat) -> TL -> AD (the symmetry test's chain,
    adjoint/validation.py:132-153) plus one perturbed_state, on the first `ncol` columns of a synthetic block,
    through `impl` = oracle.cloudsc2_numpy (the restatement) or oracle.ref_run (the reference's own stencil
    sources).  Flattened into one dict: the schema of tests/golden/ref_*.npz."""
    P = externals(**{"LREGCL": True, **flags})
    st = {k: np.ascontiguousarray(v[:, :ncol]) for k, v in make_state(block, dtype).items()}
    s = dict(st)
    s["f_eta"] = onp.eta_levels(s["f_ap"], s["f_aph"])  # common/diagnostics.py:42-45 (a python loop, not a stencil)
    s["f_qsat"] = impl.saturation(s["f_ap"], s["f_t"], P)
    out = {"in_f_eta": s["f_eta"], "in_f_qsat": s["f_qsat"]}
    tn, dg = impl.cloudsc2_nl(s, DT, P)
    out.update({"nl_t_" + k: v for k, v in tn.items()})
    out.update({"nl_d_" + k: v for k, v in dg.items()})
    si = impl.state_increment(s, 0.01, ignore_supsat=True)
    s.update(si)
    out.update({"inc_" + k: v for k, v in si.items()})
    out.update({"pert_" + k: v for k, v in impl.perturbed_state(s, 1e-3).items()})
    ttl, dtl = impl.cloudsc2_tl(s, DT, P)
    out.update({"tl_t_" + k: v.copy() for k, v in ttl.items()})
    out.update({"tl_d_" + k: v.copy() for k, v in dtl.items()})
    ad_in = dict(s)
    for x in ("t", "q", "ql", "qi"):
        ad_in[f"f_tnd_{x}_i"] = ttl[f"f_{x}_i"].copy()
    for k, v in dtl.items():
        ad_in[k] = v.copy()
    tad, dad = impl.cloudsc2_ad(ad_in, DT, P, **(ad_kwargs or {}))
    out.update({"ad_t_" + k: v for k, v in tad.items()})
    out.update({"ad_d_" + k: v for k, v in dad.items()})
    out.update({"ad_seed_" + k: ad_in[k] for k in SEED_KEYS})
    return out


# the fixtures written from the reference's own sources: name -> (block, precision, ncol, externals overrides)
REF_FIXTURES = {
    "ref_base_double": ("base", np.float64, 32, {}),
    "ref_cold_double": ("cold", np.float64, 32, {}),
    "ref_base_single": ("base", np.float32, 32, {}),
    "ref_cold_single": ("cold", np.float32, 32, {}),
    "ref_base_double_noregcl": ("base", np.float64, 16, {"LREGCL": False}),
    "ref_base_double_tetens": ("base", np.float64, 16, {"LPHYLIN": False}),
    "ref_base_double_tetens_kflag0": ("base", np.float64, 16, {"LPHYLIN": False, "KFLAG": 0}),
    "ref_base_double_levapls2": ("base", np.float64, 16, {"LEVAPLS2": True}),
    "ref_base_double_ldrain1d": ("base", np.float64, 16, {"LDRAIN1D": True}),
}


# ------------------------------------------------------------------------------------------
# host twin (csrc column code compiled for the CPU; see oracle/host_twin/twin.cpp)
# ------------------------------------------------------------------------------------------
_twin = None


def build_twin() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)


def twin():
    global _twin
    if _twin is None:
        if not os.path.exists(TWIN_PATH):
            build_twin()
        _twin = C.CDLL(TWIN_PATH)
    return _twin


def stride_of(ncol: int) -> int:
    return max(32, -(-ncol // 32) * 32)


class HostFields:
    """numpy arrays in the product layout [nlev+1][stride] + ctypes pointers to them."""

    def __init__(self, ncol: int, nlev: int, dtype):
        self.ncol, self.nlev, self.dtype = ncol, nlev, np.dtype(dtype)
        self.stride = stride_of(ncol)
        self.arrays: Dict[str, np.ndarray] = {}

    def put(self, name: str, arr: np.ndarray) -> int:
        buf = np.zeros((self.nlev + 1, self.stride), dtype=self.dtype)
        buf[:, : self.ncol] = arr
        self.arrays[name] = buf
        return buf.ctypes.data

    def new(self, name: str) -> int:
        buf = np.zeros((self.nlev + 1, self.stride), dtype=self.dtype)
        self.arrays[name] = buf
        return buf.ctypes.data

    def get(self, name: str) -> np.ndarray:
        return self.arrays[name][:, : self.ncol].copy()

    def dims(self) -> "_lib.Dims":
        return _lib.Dims(self.ncol, self.stride, self.nlev, _lib.CS2_F64 if self.dtype == np.float64 else _lib.CS2_F32)


def level_tables(P: Dict, eta: np.ndarray, nlev: int, dtype) -> np.ndarray:
    lib = _lib.load()
    code = _lib.CS2_F64 if np.dtype(dtype) == np.float64 else _lib.CS2_F32
    nbytes = lib.cs2_level_tables_bytes(nlev, code)
    buf = np.zeros(nbytes, dtype=np.uint8)
    eta_c = np.ascontiguousarray(eta[:nlev].astype(dtype))
    params = _lib.make_params(P)
    _lib.check(
        lib.cs2_level_tables_build(C.byref(params), nlev, code, eta_c.ctypes.data, buf.ctypes.data, nbytes),
        "cs2_level_tables_build",
    )
    return buf


_NL_STATE = {
    "in_ap": "f_ap", "in_aph": "f_aph", "in_lu": "f_lu", "in_lude": "f_lude", "in_mfd": "f_mfd", "in_mfu": "f_mfu",
    "in_q": "f_q", "in_qi": "f_qi", "in_ql": "f_ql", "in_qsat": "f_qsat", "in_supsat": "f_supsat", "in_t": "f_t",
    "in_tnd_cml_q": "f_tnd_cml_q", "in_tnd_cml_qi": "f_tnd_cml_qi", "in_tnd_cml_ql": "f_tnd_cml_ql",
    "in_tnd_cml_t": "f_tnd_cml_t",
}
_NL_OUT_T = {"out_tnd_q": "f_q", "out_tnd_qi": "f_qi", "out_tnd_ql": "f_ql", "out_tnd_t": "f_t"}
_NL_OUT_D = {
    "out_clc": "f_clc", "out_covptot": "f_covptot", "out_fhpsl": "f_fhpsl", "out_fhpsn": "f_fhpsn",
    "out_fplsl": "f_fplsl", "out_fplsn": "f_fplsn",
}


def _nl_struct(h: HostFields, s: Dict[str, np.ndarray], suffix: str = "") -> "_lib.NLFields":
    f = _lib.NLFields()
    for arg, key in _NL_STATE.items():
        setattr(f, arg, h.put(arg + suffix, s[key + suffix]))
    for arg in list(_NL_OUT_T) + list(_NL_OUT_D):
        setattr(f, arg, h.new(arg + suffix))
    return f


def _collect_nl(h: HostFields, suffix: str = ""):
    tends = {key + suffix: h.get(arg + suffix) for arg, key in _NL_OUT_T.items()}
    diags = {key + suffix: h.get(arg + suffix) for arg, key in _NL_OUT_D.items()}
    return tends, diags


def twin_saturation(ap, t, P):
    nlev, ncol = ap.shape[0] - 1, ap.shape[1]
    h = HostFields(ncol, nlev, ap.dtype)
    pa, pt, pq = h.put("ap", ap), h.put("t", t), h.new("qsat")
    params, dims = _lib.make_params(P), h.dims()
    twin().twin_saturation(C.byref(dims), C.byref(params), C.c_void_p(pa), C.c_void_p(pt), C.c_void_p(pq))
    return h.get("qsat")


def twin_nl(s, dt, P, split=False, pipe=False, ad_ref=False):
    """`split`: through the two half-level functions of the split NL kernel (cs2_physics_split.cuh);
    `pipe`: through the software-pipelined level function of the default NL kernel (cs2_physics_pipe.cuh)."""
    nlev, ncol = s["f_ap"].shape[0] - 1, s["f_ap"].shape[1]
    h = HostFields(ncol, nlev, s["f_ap"].dtype)
    f = _nl_struct(h, s)
    tab = level_tables(P, s["f_eta"], nlev, h.dtype)
    params, dims = _lib.make_params(P), h.dims()
    if pipe:
        rc = twin().twin_nl_pipe(C.byref(dims), C.byref(params), C.c_double(dt), C.c_void_p(tab.ctypes.data), C.byref(f),
                                 C.c_int(1 if ad_ref else 0))
    else:
        fn = twin().twin_nl_split if split else twin().twin_nl
        rc = fn(C.byref(dims), C.byref(params), C.c_double(dt), C.c_void_p(tab.ctypes.data), C.byref(f))
    assert rc == 0
    return _collect_nl(h)


def twin_tl(s, dt, P):
    nlev, ncol = s["f_ap"].shape[0] - 1, s["f_ap"].shape[1]
    h = HostFields(ncol, nlev, s["f_ap"].dtype)
    f = _nl_struct(h, s)
    g = _nl_struct(h, s, "_i")
    tab = level_tables(P, s["f_eta"], nlev, h.dtype)
    params, dims = _lib.make_params(P), h.dims()
    twin().twin_tl(C.byref(dims), C.byref(params), C.c_double(dt), C.c_void_p(tab.ctypes.data), C.byref(f), C.byref(g))
    t0, d0 = _collect_nl(h)
    t1, d1 = _collect_nl(h, "_i")
    return {**t0, **t1}, {**d0, **d1}


_AD_SEED_KEYS = {
    "in_tnd_t_i": "f_tnd_t_i", "in_tnd_q_i": "f_tnd_q_i", "in_tnd_ql_i": "f_tnd_ql_i", "in_tnd_qi_i": "f_tnd_qi_i",
    "in_clc_i": "f_clc_i", "in_covptot_i": "f_covptot_i", "in_fhpsl_i": "f_fhpsl_i", "in_fhpsn_i": "f_fhpsn_i",
    "in_fplsl_i": "f_fplsl_i", "in_fplsn_i": "f_fplsn_i",
}
_AD_OUT_T = {
    "out_tnd_cml_t_i": "f_cml_t_i", "out_tnd_cml_q_i": "f_cml_q_i", "out_tnd_cml_ql_i": "f_cml_ql_i",
    "out_tnd_cml_qi_i": "f_cml_qi_i",
}
_AD_OUT_D = {
    "out_aph_i": "f_aph_i", "out_ap_i": "f_ap_i", "out_q_i": "f_q_i", "out_qsat_i": "f_qsat_i", "out_t_i": "f_t_i",
    "out_ql_i": "f_ql_i", "out_qi_i": "f_qi_i", "out_lude_i": "f_lude_i", "out_lu_i": "f_lu_i", "out_mfu_i": "f_mfu_i",
    "out_mfd_i": "f_mfd_i", "out_supsat_i": "f_supsat_i",
}


def twin_ad(s, dt, P, predicates="tl"):
    """s: NL inputs + the 10 adjoint seeds.  Returns (tendencies, diagnostics, consumed_seeds)."""
    nlev, ncol = s["f_ap"].shape[0] - 1, s["f_ap"].shape[1]
    h = HostFields(ncol, nlev, s["f_ap"].dtype)
    f = _nl_struct(h, s)
    seeds = _lib.ADSeeds()
    for arg, key in _AD_SEED_KEYS.items():
        setattr(seeds, arg, h.put(arg, s[key]))
    outs = _lib.ADOutputs()
    for arg in list(_AD_OUT_T) + list(_AD_OUT_D):
        setattr(outs, arg, h.new(arg))
    tab = level_tables(P, s["f_eta"], nlev, h.dtype)
    params, dims = _lib.make_params(dict(P, AD_TL_PREDICATES=(predicates == "tl"))), h.dims()
    jsel = np.zeros(h.stride, dtype=np.int32)
    cov = np.zeros((nlev, h.stride), dtype=h.dtype)  # overlap carry per level (evaporation branch only)
    rc = twin().twin_ad(
        C.byref(dims), C.byref(params), C.c_double(dt), C.c_void_p(tab.ctypes.data), C.byref(f), C.byref(seeds),
        C.byref(outs), C.c_void_p(jsel.ctypes.data), C.c_void_p(cov.ctypes.data),
    )
    assert rc == 0
    tends, diags = _collect_nl(h)
    tends.update({key: h.get(arg) for arg, key in _AD_OUT_T.items()})
    diags.update({key: h.get(arg) for arg, key in _AD_OUT_D.items()})
    consumed = {key: h.get(arg) for arg, key in _AD_SEED_KEYS.items()}
    return tends, diags, consumed
