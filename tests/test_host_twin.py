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


@pytest.mark.parametrize("block", ["base", "cold"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ad_ref", [False, True])
def test_twin_nl_pipe(block, dtype, ad_ref):
    """The software-pipelined level function (csrc/experiments/cs2_physics_pipe.cuh: half B of level k together with
    half A of level k+1, branch-free) gives NL; `ad_ref` = the literal AD stencil's second freezing test."""
    P = H.externals(LREGCL=True)
    st = H.make_state(block, dtype, 257)
    s = H.with_diagnostics(st, P)
    tol = H.TOL[np.dtype(dtype)]
    ttn, tdg = H.twin_nl(s, H.DT, P, pipe=True, ad_ref=ad_ref)
    if not ad_ref:
        tn, dg = H.onp.cloudsc2_nl(s, H.DT, P)
    else:  # the trajectory outputs of the literal AD stencil (adjoint/_stencils/cloudsc2.py:146-475)
        _, _, _, ref = H.oracle_symmetry(st, P, predicates="reference")
        tn = {k: ref["tends_ad"][k] for k in ("f_t", "f_q", "f_ql", "f_qi")}
        dg = {k: ref["diags_ad"][k] for k in ("f_clc", "f_covptot", "f_fhpsl", "f_fhpsn", "f_fplsl", "f_fplsn")}
    H.assert_fields_close(ttn, tn, tol, "NL pipe tendencies: ")
    H.assert_fields_close(tdg, dg, tol, "NL pipe diagnostics: ")
