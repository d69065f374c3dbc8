"""CPU: the C-ABI library loads, exports every symbol include/cloudsc2_b200.h declares, validates
its arguments, and builds level tables identical to the oracle's per-level formulas.  No compute
entry point is launched here (there is no GPU in the CPU test tier)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from cloudsc2_b200 import _lib

HEADER = os.path.join(H.ROOT, "include", "cloudsc2_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cs2_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15, names
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype"
    assert lib.cs2_abi_version() == 1


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Dims) == 24
    assert C.sizeof(_lib.NLFields) == 26 * 8
    assert C.sizeof(_lib.ADSeeds) == 10 * 8 and C.sizeof(_lib.ADOutputs) == 16 * 8
    assert C.sizeof(_lib.Params) == 35 * 8 + 8 * 4
    text = open(HEADER).read()
    body = text[text.index("typedef struct cs2_params {") : text.index("} cs2_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    doubles = [n.strip() for line in re.findall(r"double ([^;]+);", body) for n in line.split(",")]
    ints = [n.strip() for line in re.findall(r"int32_t ([^;]+);", body) for n in line.split(",")]
    assert doubles + ints == [n for n, _ in _lib.Params._fields_]
    for struct, cname in ((_lib.NLFields, "cs2_nl_fields"), (_lib.ADSeeds, "cs2_ad_seeds"), (_lib.ADOutputs, "cs2_ad_outputs")):
        b = text[text.index(f"typedef struct {cname} {{") : text.index(f"}} {cname};")]
        members = [m.strip().lstrip("*") for line in re.findall(r"void ([^;]+);", b) for m in line.split(",")]
        assert members == [n for n, _ in struct._fields_], cname


def test_argument_validation_without_gpu():
    lib = _lib.load()
    params = _lib.make_params(H.externals())
    buf = np.zeros(64, dtype=np.float64)
    p = buf.ctypes.data
    bad_stride = _lib.Dims(10, 33, 4, _lib.CS2_F64)
    assert lib.cs2_saturation(C.byref(bad_stride), C.byref(params), p, p, p, None) == -3  # CS2_ERR_MISALIGNED
    assert b"multiple of 32" in lib.cs2_last_error()
    bad_dtype = _lib.Dims(10, 32, 4, 7)
    assert lib.cs2_saturation(C.byref(bad_dtype), C.byref(params), p, p, p, None) == -1  # CS2_ERR_BAD_DIMS
    ok = _lib.Dims(10, 32, 4, _lib.CS2_F64)
    assert lib.cs2_saturation(C.byref(ok), C.byref(params), None, p, p, None) == -2  # CS2_ERR_NULL_POINTER
    assert lib.cs2_saturation(C.byref(ok), C.byref(params), p + 8, p, p, None) == -3
    f = _lib.NLFields()
    assert lib.cs2_nl(C.byref(ok), C.byref(params), 3600.0, p, C.byref(f), None) == -2
    evap = _lib.make_params(H.externals(LEVAPLS2=True))
    for name, _ in _lib.NLFields._fields_:
        setattr(f, name, p)
    # every flag combination is implemented (NL, TL and AD incl. the evaporation branch): the workspace of AD grows by one
    # plane with LEVAPLS2 / LDRAIN1D (overlap carry per level) and the checkpoint planes are not used then
    plain = _lib.make_params(H.externals())
    big = _lib.Dims(64, 64, 137, _lib.CS2_F64)
    w_rec = lib.cs2_ad_workspace_bytes(C.byref(big), C.byref(plain), _lib.CS2_AD_RECOMPUTE)
    w_ck = lib.cs2_ad_workspace_bytes(C.byref(big), C.byref(plain), _lib.CS2_AD_CHECKPOINT)
    w_ev = lib.cs2_ad_workspace_bytes(C.byref(big), C.byref(evap), _lib.CS2_AD_CHECKPOINT)
    assert w_rec == 256 and w_ck == 256 + 9 * 137 * 64 * 8 and w_ev == 256 + 137 * 64 * 8
    empty = _lib.Dims(0, 32, 4, _lib.CS2_F64)  # zero columns: nothing to launch, no GPU needed
    assert lib.cs2_saturation(C.byref(empty), C.byref(params), p, p, p, None) == 0


def test_no_gpu_fails_loudly_not_silently():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.cs2_device_count() == 0
    params = _lib.make_params(H.externals())
    buf = np.zeros(64 * 5, dtype=np.float64)
    ok = _lib.Dims(10, 32, 4, _lib.CS2_F64)
    rc = lib.cs2_saturation(C.byref(ok), C.byref(params), buf.ctypes.data, buf.ctypes.data, buf.ctypes.data, None)
    assert rc == -5 and lib.cs2_last_error()  # CS2_ERR_CUDA


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_level_tables_match_oracle_formulas(dtype):
    """scalm and crh2 of every (level, tropopause candidate) equal the oracle's per-point values."""
    P = H.externals()
    s = H.with_diagnostics(H.make_state("base", dtype), P)
    eta, nlev = s["f_eta"], 137
    tab = H.level_tables(P, eta, nlev, dtype)
    hdr = tab[:8].view(np.int32)
    assert hdr[0] == nlev
    nw = int(hdr[1])
    es = np.dtype(dtype).itemsize
    a16 = lambda n: (n + 15) & ~15  # noqa: E731
    off = 16
    scalm = tab[off : off + nlev * es].view(dtype)
    off += a16(nlev * es)
    crh2 = tab[off : off + nlev * (nw + 1) * es].view(dtype).reshape(nlev, nw + 1)
    off += a16(nlev * (nw + 1) * es)
    wlev = tab[off : off + nw * 4].view(np.int32)
    expect_w = [k for k in range(nlev - 1) if 0.1 < eta[k] < 0.4]
    assert list(wlev) == expect_w and nw > 5
    dt = np.dtype(dtype).type
    ref_scalm = np.array([dt(P["ZSCAL"]) * max(eta[k] - dt(0.2), dt(P["ZEPS1"])) ** dt(0.2) for k in range(nlev)], dtype=dtype)
    np.testing.assert_allclose(scalm, ref_scalm, rtol=4 * np.finfo(dtype).eps)
    cands = np.array([0.1] + [eta[k] for k in wlev], dtype=dtype)
    for k in range(nlev):
        ref = H.onp._crh2(eta[k], cands, np.dtype(dtype))
        assert np.array_equal(crh2[k], ref), k
