"""ctypes binding of `libcloudsc2_b200.so` (the C ABI in include/cloudsc2_b200.h).

There is no fallback: if the shared library has not been built (`python -m cloudsc2_b200.build`
or `__graft_entry__.build()`), loading raises `CUDAExtensionError`; if it is built but no CUDA
device is present, every compute entry point returns CS2_ERR_CUDA, which `check()` turns into a
`CUDAExtensionError` as well.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Dict, Optional

LIB_NAME = "libcloudsc2_b200.so"
# CS2_LIB lets a developer A/B-test another build of the same CUDA library (tests/kbench.py); it never selects a
# different implementation: the file must export the same C ABI.
LIB_PATH = os.environ.get("CS2_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)

CS2_F64, CS2_F32 = 0, 1
CS2_AD_RECOMPUTE, CS2_AD_CHECKPOINT = 0, 1
CS2_NSTATE = 16

# order of the 16-pointer arrays of cs2_state_increment / cs2_perturbed_state
STATE_ORDER = (
    "aph", "ap", "q", "qsat", "t", "ql", "qi", "lude", "lu", "mfu", "mfd",
    "tnd_cml_t", "tnd_cml_q", "tnd_cml_ql", "tnd_cml_qi", "supsat",
)


class CUDAExtensionError(RuntimeError):
    """The CUDA extension is missing, or a call into it failed."""


class Dims(C.Structure):
    _fields_ = [("ncol", C.c_int64), ("ncol_stride", C.c_int64), ("nlev", C.c_int32), ("dtype", C.c_int32)]


_PARAM_DOUBLES = (
    "R2ES R3IES R3LES R4IES R4LES R5ALSCP R5ALVCP R5IES R5LES RALSDCP RALVDCP RTICE RTICECU RTWAT "
    "RTWAT_RTICE_R RTWAT_RTICECU_R RVTMP2 RCPD RD RETV RG RLMLT RLSTT RLVTT RTT RCLCRIT RKCONV RLMIN "
    "RPECONS RLPTRC ZEPS1 ZEPS2 ZQMAX ZSCAL QMAX"
).split()
_PARAM_INTS = "LPHYLIN LDRAIN1D LEVAPLS2 LREGCL KFLAG ICALL AD_TL_PREDICATES reserved_".split()


class Params(C.Structure):
    _fields_ = [(n, C.c_double) for n in _PARAM_DOUBLES] + [(n, C.c_int32) for n in _PARAM_INTS]


_NL_IN = (
    "in_ap in_aph in_lu in_lude in_mfd in_mfu in_q in_qi in_ql in_qsat in_supsat in_t in_tnd_cml_q "
    "in_tnd_cml_qi in_tnd_cml_ql in_tnd_cml_t"
).split()
_NL_OUT = "out_clc out_covptot out_fhpsl out_fhpsn out_fplsl out_fplsn out_tnd_q out_tnd_qi out_tnd_ql out_tnd_t".split()


class NLFields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _NL_IN + _NL_OUT]


_AD_SEEDS = (
    "in_tnd_t_i in_tnd_q_i in_tnd_ql_i in_tnd_qi_i in_clc_i in_covptot_i in_fhpsl_i in_fhpsn_i in_fplsl_i in_fplsn_i"
).split()
_AD_OUT = (
    "out_aph_i out_ap_i out_q_i out_qsat_i out_t_i out_ql_i out_qi_i out_lude_i out_lu_i out_mfu_i out_mfd_i "
    "out_supsat_i out_tnd_cml_t_i out_tnd_cml_q_i out_tnd_cml_ql_i out_tnd_cml_qi_i"
).split()


class ADSeeds(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _AD_SEEDS]


class ADOutputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _AD_OUT]


NL_IN_NAMES, NL_OUT_NAMES = tuple(_NL_IN), tuple(_NL_OUT)
AD_SEED_NAMES, AD_OUT_NAMES = tuple(_AD_SEEDS), tuple(_AD_OUT)

PtrArray16 = C.c_void_p * CS2_NSTATE

# every symbol include/cloudsc2_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "cs2_abi_version": (C.c_int, []),
    "cs2_last_error": (C.c_char_p, []),
    "cs2_device_count": (C.c_int, []),
    "cs2_dfma_rate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "cs2_dfma_rate_regs": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "cs2_level_tables_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "cs2_level_tables_build": (C.c_int, [C.POINTER(Params), C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "cs2_saturation": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cs2_state_increment": (C.c_int, [C.POINTER(Dims), C.c_double, C.c_int32, PtrArray16, PtrArray16, C.c_void_p]),
    "cs2_perturbed_state": (C.c_int, [C.POINTER(Dims), C.c_double, PtrArray16, PtrArray16, PtrArray16, C.c_void_p]),
    "cs2_nl": (C.c_int, [C.POINTER(Dims), C.POINTER(Params), C.c_double, C.c_void_p, C.POINTER(NLFields), C.c_void_p]),
    "cs2_tl_increment": (
        C.c_int,
        [C.POINTER(Dims), C.POINTER(Params), C.c_double, C.c_void_p, C.POINTER(NLFields), C.POINTER(NLFields), C.c_double,
         C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "cs2_nl_perturbed": (
        C.c_int,
        [C.POINTER(Dims), C.POINTER(Params), C.c_double, C.c_void_p, C.POINTER(NLFields), C.POINTER(NLFields), C.c_double,
         C.c_void_p],
    ),
    "cs2_tl": (
        C.c_int,
        [C.POINTER(Dims), C.POINTER(Params), C.c_double, C.c_void_p, C.POINTER(NLFields), C.POINTER(NLFields), C.c_void_p],
    ),
    "cs2_ad_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.POINTER(Params), C.c_int32]),
    "cs2_ad": (
        C.c_int,
        [
            C.POINTER(Dims), C.POINTER(Params), C.c_double, C.c_void_p, C.POINTER(NLFields), C.POINTER(ADSeeds),
            C.POINTER(ADOutputs), C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p,
        ],
    ),
    "cs2_ad_norm2": (
        C.c_int,
        [
            C.POINTER(Dims), C.POINTER(Params), C.c_double, C.c_void_p, C.POINTER(NLFields), C.POINTER(ADSeeds),
            C.POINTER(ADOutputs), C.c_void_p, C.c_size_t, C.c_int32, C.c_double, C.c_int32, C.c_void_p, C.c_void_p,
        ],
    ),
    "cs2_taylor_scratch_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int32]),
    "cs2_taylor_sums": (
        C.c_int,
        [
            C.POINTER(Dims), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
            C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
        ],
    ),
    "cs2_taylor_nl_scratch_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "cs2_taylor_nl_sums": (
        C.c_int,
        [C.POINTER(Dims), C.POINTER(Params), C.c_double, C.c_void_p, C.POINTER(NLFields), C.c_double, C.c_int32, C.c_double,
         C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "cs2_symmetry_norms": (
        C.c_int,
        [C.POINTER(Dims), C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p],
    ),
    "cs2_symmetry_residual_scratch_bytes": (C.c_size_t, [C.c_int64]),
    "cs2_symmetry_residual": (
        C.c_int,
        [C.c_int64, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the prototypes.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CUDAExtensionError(
            f"{LIB_PATH} not found: the CUDA extension has not been built.  Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m cloudsc2_b200.build`) first. "
            "There is no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the machine
        raise CUDAExtensionError(f"cannot load {LIB_PATH}: {exc}") from exc
    for name, (restype, argtypes) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as exc:
            raise CUDAExtensionError(f"{LIB_PATH} does not export {name}") from exc
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.cs2_abi_version() != 1:
        raise CUDAExtensionError(f"{LIB_PATH}: ABI version {lib.cs2_abi_version()} != 1 (stale build?)")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cs2_last_error().decode("utf-8", "replace")
        raise CUDAExtensionError(f"{what} failed with code {rc}: {msg}")


def make_params(externals: Dict[str, Any]) -> Params:
    """Build the `cs2_params` struct from a reference-style externals dict (missing members
    that a given stencil does not read default to 0)."""
    p = Params()
    for n in _PARAM_DOUBLES:
        setattr(p, n, float(externals.get(n, 0.0)))
    for n in _PARAM_INTS:
        setattr(p, n, int(bool(externals.get(n, 0))) if n not in ("KFLAG", "ICALL") else int(externals.get(n, 0)))
    if "QMAX" not in externals:
        p.QMAX = 0.5
    return p
