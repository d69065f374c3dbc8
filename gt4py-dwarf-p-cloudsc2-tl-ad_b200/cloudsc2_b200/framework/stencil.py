"""Stencil registry: `compile_stencil(name, externals)` returns a callable with the keyword
signature of the corresponding GT4Py StencilObject of the reference, backed by the C ABI of
`libcloudsc2_b200.so` (include/cloudsc2_b200.h).  This is the plug-in boundary of the reference
(SURVEY.md section 8b): `array_call` bodies stay as in the reference, the object they call is
replaced.

Field arguments are the logical `(nx, 1, nz+1)` torch views handed out by `Field.data`
(column-fastest storage, see storage.py).  They must live on a CUDA device; there is no CPU
fallback -- calling a stencil with host tensors raises `CUDAExtensionError`.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Any, Callable, Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from .._lib import CUDAExtensionError

_REGISTRY: Dict[str, Callable[..., "StencilObject"]] = {}


def stencil_collection(name: str):
    def deco(cls):
        _REGISTRY[name] = cls
        cls.stencil_name = name
        return cls

    return deco


def compile_stencil(name: str, externals: Optional[Dict[str, Any]] = None, gt4py_config: Any = None) -> "StencilObject":
    if name not in _REGISTRY:
        raise KeyError(f"unknown stencil {name!r}; available: {sorted(_REGISTRY)}")
    return _REGISTRY[name](externals or {}, gt4py_config)


class StencilObject:
    stencil_name = ""

    def __init__(self, externals: Dict[str, Any], gt4py_config: Any = None) -> None:
        self.externals = dict(externals)
        self.gt4py_config = gt4py_config
        self.lib = _lib.load()  # fails loudly when the extension is not built
        self.params = _lib.make_params(self.externals)
        self._tables_key = None
        self._tables_dev: Optional[torch.Tensor] = None
        self._events = []
        # validated (layout, device pointer) per tensor OBJECT: id -> (weakref, (nx, nlevs, stride, code, ptr)).  The same
        # views come back call after call (Field.data hands out one view object per field), so the layout checks run once.
        self._ptr_cache: Dict[int, Tuple[Any, Tuple[int, int, int, int, int]]] = {}

    # ---- helpers -----------------------------------------------------------------------
    @staticmethod
    def _layout(arr: torch.Tensor, what: str) -> Tuple[int, int, int, int]:
        """(nx, nlevs, stride, dtype code) of a logical (nx, 1, nlevs) view; validates the layout."""
        if not isinstance(arr, torch.Tensor):
            raise TypeError(f"{what}: expected a torch.Tensor view of a cloudsc2_b200 Field, got {type(arr).__name__}")
        if arr.device.type != "cuda":
            raise CUDAExtensionError(
                f"{what}: field lives on {arr.device}; the CLOUDSC2 stencils run only as CUDA kernels on a B200 "
                "(there is no CPU fallback)"
            )
        if arr.dim() != 3 or arr.shape[1] != 1:
            raise ValueError(f"{what}: expected logical shape (nx, 1, nz+1), got {tuple(arr.shape)}")
        sx, _, sk = arr.stride()
        if (arr.shape[0] > 1 and sx != 1) or sk % 32 != 0:
            raise ValueError(
                f"{what}: field is not in column-fastest storage (strides {arr.stride()}); allocate it with "
                "cloudsc2_b200.framework.storage.zeros()"
            )
        if arr.dtype == torch.float64:
            code = _lib.CS2_F64
        elif arr.dtype == torch.float32:
            code = _lib.CS2_F32
        else:
            raise TypeError(f"{what}: dtype {arr.dtype} is not float32/float64")
        return arr.shape[0], arr.shape[2], sk, code

    def _dims(self, ref: torch.Tensor, what: str, nlev: Optional[int] = None) -> _lib.Dims:
        nx, nlevs, stride, code = self._layout(ref, what)
        return _lib.Dims(nx, stride, (nlevs - 1) if nlev is None else nlev, code)

    def _ptr(self, arr: torch.Tensor, dims: _lib.Dims, what: str) -> int:
        hit = self._ptr_cache.get(id(arr))
        if hit is not None and hit[0]() is arr:
            nx, nlevs, stride, code, ptr = hit[1]
        else:
            nx, nlevs, stride, code = self._layout(arr, what)
            # (torch reports data_ptr() == 0 for views without elements, e.g. an empty column shard)
            ptr = arr.untyped_storage().data_ptr() + arr.storage_offset() * arr.element_size()
            if len(self._ptr_cache) >= 1024:  # entries of tensors that no longer exist
                self._ptr_cache = {k: v for k, v in self._ptr_cache.items() if v[0]() is not None}
                if len(self._ptr_cache) >= 1024:
                    self._ptr_cache.clear()
            self._ptr_cache[id(arr)] = (weakref.ref(arr), (nx, nlevs, stride, code, ptr))
        if nx != dims.ncol or stride != dims.ncol_stride or code != dims.dtype or nlevs != dims.nlev + 1:
            raise ValueError(f"{what}: layout {(nx, nlevs, stride, code)} differs from the call's "
                             f"{(dims.ncol, dims.nlev + 1, dims.ncol_stride, dims.dtype)}")
        return ptr

    @staticmethod
    def _stream(ref: torch.Tensor) -> int:
        return torch.cuda.current_stream(ref.device).cuda_stream

    def _level_tables(self, eta: Any, nlev: int, dims: _lib.Dims, device: torch.device) -> torch.Tensor:
        np_dtype = np.float64 if dims.dtype == _lib.CS2_F64 else np.float32
        eta_np = eta.detach().cpu().numpy() if isinstance(eta, torch.Tensor) else np.asarray(eta)
        eta_np = np.ascontiguousarray(eta_np.reshape(-1)[:nlev].astype(np_dtype))
        key = (eta_np.tobytes(), nlev, dims.dtype, str(device), bytes(self.params))
        if key != self._tables_key:
            nbytes = self.lib.cs2_level_tables_bytes(nlev, dims.dtype)
            host = torch.zeros(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            _lib.check(
                self.lib.cs2_level_tables_build(C.byref(self.params), nlev, dims.dtype, eta_np.ctypes.data,
                                                host.data_ptr(), nbytes),
                "cs2_level_tables_build",
            )
            self._tables_dev = host.to(device)
            self._tables_key = key
        return self._tables_dev

    def _check_domain(self, domain: Optional[Tuple[int, ...]], nx: int, nk: int) -> None:
        if domain is not None and (int(domain[0]) != nx or int(domain[2]) != nk):
            raise ValueError(f"{self.stencil_name}: domain {tuple(domain)} does not match the fields ({nx}, 1, {nk})")

    class _Timer:
        def __init__(self, owner: "StencilObject", exec_info: Optional[Dict[str, Any]], device: torch.device) -> None:
            self.owner, self.exec_info, self.device = owner, exec_info, device

        def __enter__(self):
            if self.exec_info is not None:
                self.start = torch.cuda.Event(enable_timing=True)
                self.stop = torch.cuda.Event(enable_timing=True)
                self.start.record(torch.cuda.current_stream(self.device))
            return self

        def __exit__(self, *exc):
            if self.exec_info is not None and exc[0] is None:
                self.stop.record(torch.cuda.current_stream(self.device))
                rec = self.exec_info.setdefault(self.owner.stencil_name, {"ncalls": 0, "_pending": [], "total_run_time": 0.0})
                rec["ncalls"] += 1
                rec["_pending"].append((self.start, self.stop))
            return False


def resolve_exec_info(exec_info: Optional[Dict[str, Any]]) -> Dict[str, Dict[str, float]]:
    """Turn the pending CUDA event pairs recorded by the stencil calls into seconds
    (the analogue of GT4Py's `exec_info["<stencil>"]["total_run_time"]`)."""
    out: Dict[str, Dict[str, float]] = {}
    if not exec_info:
        return out
    torch.cuda.synchronize()
    for name, rec in exec_info.items():
        for start, stop in rec.pop("_pending", []):
            rec["total_run_time"] += start.elapsed_time(stop) * 1e-3
        rec["_pending"] = []
        out[name] = {"ncalls": rec["ncalls"], "total_run_time": rec["total_run_time"]}
    return out


# ----------------------------------------------------------------------------------------------
@stencil_collection("saturation")
class SaturationStencil(StencilObject):
    """common/_stencils/saturation.py:23-42 -> cs2_saturation"""

    def __call__(self, *, in_ap, in_t, out_qsat, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None):
        nlev = int(domain[2]) if domain is not None else in_ap.shape[2] - 1
        dims = self._dims(in_ap, "in_ap")
        if nlev > dims.nlev + 1:
            raise ValueError("saturation: domain has more levels than the storage")
        self._check_domain(domain, dims.ncol, nlev)
        call_dims = _lib.Dims(dims.ncol, dims.ncol_stride, nlev, dims.dtype)
        ptrs = [self._ptr(a, dims, n) for a, n in ((in_ap, "in_ap"), (in_t, "in_t"), (out_qsat, "out_qsat"))]
        with self._Timer(self, exec_info, in_ap.device):
            _lib.check(self.lib.cs2_saturation(C.byref(call_dims), C.byref(self.params), ptrs[0], ptrs[1], ptrs[2],
                                               self._stream(in_ap)), "cs2_saturation")


def _state_array(kwargs: Dict[str, Any], prefix: str, suffix: str, dims: _lib.Dims, obj: StencilObject):
    arr = _lib.PtrArray16()
    for n, name in enumerate(_lib.STATE_ORDER):
        key = f"{prefix}{name}{suffix}"
        arr[n] = obj._ptr(kwargs[key], dims, key)
    return arr


@stencil_collection("state_increment")
class StateIncrementStencil(StencilObject):
    """common/_stencils/state_increment.py:22-80 -> cs2_state_increment"""

    def __call__(self, *, f, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None, **fields):
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        ins = _state_array(fields, "in_", "", dims, self)
        outs = _state_array(fields, "out_", "_i", dims, self)
        with self._Timer(self, exec_info, ref.device):
            _lib.check(self.lib.cs2_state_increment(C.byref(dims), float(f), int(bool(self.externals.get("IGNORE_SUPSAT", False))),
                                                    ins, outs, self._stream(ref)), "cs2_state_increment")


@stencil_collection("perturbed_state")
class PerturbedStateStencil(StencilObject):
    """common/_stencils/perturbed_state.py:22-91 -> cs2_perturbed_state"""

    def __call__(self, *, f, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None, **fields):
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        ins = _state_array(fields, "in_", "", dims, self)
        ins_i = _state_array(fields, "in_", "_i", dims, self)
        outs = _state_array(fields, "out_", "", dims, self)
        with self._Timer(self, exec_info, ref.device):
            _lib.check(self.lib.cs2_perturbed_state(C.byref(dims), float(f), ins, ins_i, outs, self._stream(ref)),
                       "cs2_perturbed_state")


def _nl_struct(obj: StencilObject, fields: Dict[str, Any], dims: _lib.Dims, suffix: str = "") -> _lib.NLFields:
    s = _lib.NLFields()
    for name in _lib.NL_IN_NAMES + _lib.NL_OUT_NAMES:
        key = name + suffix
        setattr(s, name, obj._ptr(fields[key], dims, key))
    return s


@stencil_collection("cloudsc2_nl")
class Cloudsc2NLStencil(StencilObject):
    """nonlinear/_stencils/cloudsc2.py:24-399 -> cs2_nl"""

    def __call__(self, *, in_eta, dt, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None, **fields):
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        f = _nl_struct(self, fields, dims)
        tables = self._level_tables(in_eta, dims.nlev, dims, ref.device)
        with self._Timer(self, exec_info, ref.device):
            _lib.check(self.lib.cs2_nl(C.byref(dims), C.byref(self.params), float(dt), tables.data_ptr(), C.byref(f),
                                       self._stream(ref)), "cs2_nl")


@stencil_collection("cloudsc2_nl_perturbed")
class Cloudsc2NLPerturbedStencil(StencilObject):
    """Fusion of `perturbed_state` and `cloudsc2_nl` (no counterpart stencil in the reference; it computes what
    tangent_linear/validation.py:167-176 computes with two stencil calls) -> cs2_nl_perturbed"""

    def __call__(self, *, in_eta, dt, f, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None, **fields):
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        base = _nl_struct(self, fields, dims)
        incr = _lib.NLFields()
        for name in _lib.NL_IN_NAMES:
            setattr(incr, name, self._ptr(fields[name + "_i"], dims, name + "_i"))
        tables = self._level_tables(in_eta, dims.nlev, dims, ref.device)
        with self._Timer(self, exec_info, ref.device):
            _lib.check(self.lib.cs2_nl_perturbed(C.byref(dims), C.byref(self.params), float(dt), tables.data_ptr(),
                                                 C.byref(base), C.byref(incr), float(f), self._stream(ref)),
                       "cs2_nl_perturbed")


@stencil_collection("cloudsc2_nl_taylor_sums")
class Cloudsc2NLTaylorSumsStencil(StencilObject):
    """One factor of the Taylor test in one sweep: `state_increment(f1)` -> `perturbed_state(f2)` -> `cloudsc2_nl` and the
    field sums of get_field_norm (tangent_linear/validation.py:158-176,252-261) -> cs2_taylor_nl_sums.  `in_*` = base state,
    `out_*` = the UNPERTURBED NL outputs (read); `sums` = fp64 device tensor [10][2], SUM(F_p - F_nl) is added to [:, 0]."""

    def __init__(self, externals: Dict[str, Any], gt4py_config: Any = None) -> None:
        super().__init__(externals, gt4py_config)
        self._scratch: Optional[torch.Tensor] = None

    def __call__(self, *, in_eta, dt, f1, f2, sums, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None,
                 **fields):
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        f = _nl_struct(self, fields, dims)
        tables = self._level_tables(in_eta, dims.nlev, dims, ref.device)
        nbytes = max(16, self.lib.cs2_taylor_nl_scratch_bytes(C.byref(dims)))
        if self._scratch is None or self._scratch.numel() < nbytes or self._scratch.device != ref.device:
            self._scratch = torch.empty(nbytes, dtype=torch.uint8, device=ref.device)
        assert sums.dtype == torch.float64 and sums.numel() >= 20 and sums.device == ref.device and sums.is_contiguous()
        with self._Timer(self, exec_info, ref.device):
            _lib.check(
                self.lib.cs2_taylor_nl_sums(C.byref(dims), C.byref(self.params), float(dt), tables.data_ptr(), C.byref(f),
                                            float(f1), int(bool(self.externals.get("IGNORE_SUPSAT", False))), float(f2),
                                            sums.data_ptr(), self._scratch.data_ptr(), self._scratch.numel(),
                                            self._stream(ref)),
                "cs2_taylor_nl_sums",
            )


@stencil_collection("cloudsc2_tl")
class Cloudsc2TLStencil(StencilObject):
    """tangent_linear/_stencils/cloudsc2.py:23-774 -> cs2_tl"""

    def __call__(self, *, in_eta, dt, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None, **fields):
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        f = _nl_struct(self, fields, dims)
        g = _nl_struct(self, fields, dims, "_i")
        tables = self._level_tables(in_eta, dims.nlev, dims, ref.device)
        with self._Timer(self, exec_info, ref.device):
            _lib.check(self.lib.cs2_tl(C.byref(dims), C.byref(self.params), float(dt), tables.data_ptr(), C.byref(f),
                                       C.byref(g), self._stream(ref)), "cs2_tl")


@stencil_collection("cloudsc2_tl_increment")
class Cloudsc2TLIncrementStencil(StencilObject):
    """Fusion of `state_increment` and `cloudsc2_tl` (tangent_linear/validation.py:158-162 and
    adjoint/validation.py:136-140 run them back to back) -> cs2_tl_increment.  Takes the trajectory inputs `in_*`,
    the factor `f` and both output sets; the `in_*_i` perturbations are formed in the kernel as f * in_*."""

    def __call__(self, *, in_eta, dt, f, norm1=None, origin=(0, 0, 0), domain=None, validate_args=False, exec_info=None,
                 **fields):
        """`norm1`: optional fp64 device tensor [ncol]; receives SUM_k SUM_fields (TL output)^2 per column."""
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        if norm1 is not None:
            assert norm1.dtype == torch.float64 and norm1.numel() >= dims.ncol and norm1.device == ref.device
        traj = _nl_struct(self, fields, dims)
        pert = _lib.NLFields()
        for name in _lib.NL_OUT_NAMES:
            setattr(pert, name, self._ptr(fields[name + "_i"], dims, name + "_i"))
        tables = self._level_tables(in_eta, dims.nlev, dims, ref.device)
        with self._Timer(self, exec_info, ref.device):
            _lib.check(
                self.lib.cs2_tl_increment(C.byref(dims), C.byref(self.params), float(dt), tables.data_ptr(), C.byref(traj),
                                          C.byref(pert), float(f), int(bool(self.externals.get("IGNORE_SUPSAT", False))),
                                          norm1.data_ptr() if norm1 is not None and norm1.numel() else None,
                                          self._stream(ref)),
                "cs2_tl_increment",
            )


@stencil_collection("cloudsc2_ad")
class Cloudsc2ADStencil(StencilObject):
    """adjoint/_stencils/cloudsc2.py:24-996 -> cs2_ad"""

    def __init__(self, externals: Dict[str, Any], gt4py_config: Any = None) -> None:
        super().__init__(externals, gt4py_config)
        # trajectory handling of the backward sweep: "auto" (default), "checkpoint"
        # or "recompute" (no workspace beyond 4 bytes per column)
        mode = str(externals.get("AD_TRAJECTORY", "auto"))
        if mode not in ("recompute", "checkpoint", "auto"):
            raise ValueError("AD_TRAJECTORY must be 'recompute', 'checkpoint' or 'auto'")
        # "auto": checkpoint up to AUTO_RECOMPUTE_COLUMNS columns per call, recompute above.  Measured on B200 with the
        # lockstep exponentials and the in-kernel seed reset: recompute 1.19 vs checkpoint 1.21 ms at 65 536 columns,
        # 16.2 vs 17.9 ms at 1 048 576 (where the checkpoints would also take 10 GB) -> the threshold is 0.
        self._mode_name = mode
        self.mode = _lib.CS2_AD_CHECKPOINT if mode == "checkpoint" else _lib.CS2_AD_RECOMPUTE
        self._workspace: Optional[torch.Tensor] = None

    AUTO_RECOMPUTE_COLUMNS = 0

    def _resolve_mode(self, ncol: int) -> int:
        if self._mode_name == "auto":
            return _lib.CS2_AD_CHECKPOINT if ncol <= self.AUTO_RECOMPUTE_COLUMNS else _lib.CS2_AD_RECOMPUTE
        return self.mode

    def __call__(self, *, in_eta, dt, norm2=None, increment_factor=None, origin=(0, 0, 0), domain=None, validate_args=False,
                 exec_info=None, **fields):
        """`norm2` (+ `increment_factor`; IGNORE_SUPSAT from the externals): optional fp64 device tensor [ncol]; receives
        SUM_k SUM_fields (factor * input) * (adjoint output) per column from the backward sweep (cs2_ad_norm2)."""
        ref = fields["in_ap"]
        dims = self._dims(ref, "in_ap")
        self._check_domain(domain, dims.ncol, dims.nlev + 1)
        f = _nl_struct(self, fields, dims)
        seeds = _lib.ADSeeds()
        for name in _lib.AD_SEED_NAMES:
            setattr(seeds, name, self._ptr(fields[name], dims, name))
        outs = _lib.ADOutputs()
        for name in _lib.AD_OUT_NAMES:
            setattr(outs, name, self._ptr(fields[name], dims, name))
        tables = self._level_tables(in_eta, dims.nlev, dims, ref.device)
        self.mode = self._resolve_mode(dims.ncol)
        nbytes = self.lib.cs2_ad_workspace_bytes(C.byref(dims), C.byref(self.params), self.mode)
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != ref.device:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=ref.device)
        with self._Timer(self, exec_info, ref.device):
            if norm2 is not None and dims.ncol > 0:
                assert norm2.dtype == torch.float64 and norm2.numel() >= dims.ncol and norm2.device == ref.device
                _lib.check(
                    self.lib.cs2_ad_norm2(C.byref(dims), C.byref(self.params), float(dt), tables.data_ptr(), C.byref(f),
                                          C.byref(seeds), C.byref(outs), self._workspace.data_ptr(), self._workspace.numel(),
                                          self.mode, float(increment_factor),
                                          int(bool(self.externals.get("IGNORE_SUPSAT", False))), norm2.data_ptr(),
                                          self._stream(ref)),
                    "cs2_ad_norm2",
                )
                return
            _lib.check(
                self.lib.cs2_ad(C.byref(dims), C.byref(self.params), float(dt), tables.data_ptr(), C.byref(f),
                                C.byref(seeds), C.byref(outs), self._workspace.data_ptr(), self._workspace.numel(),
                                self.mode, self._stream(ref)),
                "cs2_ad",
            )
