"""Component base classes (stand-ins for `ifs_physics_common.components`, sympl-style).

    DiagnosticComponent:        comp(state) / comp(state, out=dict)                -> diagnostics
    ImplicitTendencyComponent:  comp(state, timestep) /
                                comp(state, timestep, out_tendencies=, out_diagnostics=)
                                                                                   -> (tendencies, diagnostics)

as used by the reference drivers (drivers/run_nonlinear.py:93,109,117-118).  `state` maps field
names to `Field` objects (plus a "time" key); outputs are allocated on first use and re-used
when passed back through `out*`.  `enable_checks` validates presence and grid dims of inputs.
"""
from __future__ import annotations

from datetime import timedelta
from typing import Any, Dict, Optional, Tuple

from .config import GT4PyConfig
from .grid import ComputationalGrid
from .stencil import StencilObject, compile_stencil
from .storage import Field, allocate_field

PropertyDict = Dict[str, Dict[str, Any]]


class _Component:
    def __init__(self, computational_grid: ComputationalGrid, *, enable_checks: bool = True, gt4py_config: GT4PyConfig) -> None:
        self.computational_grid = computational_grid
        self.enable_checks = enable_checks
        self.gt4py_config = gt4py_config

    def compile_stencil(self, name: str, externals: Optional[Dict[str, Any]] = None) -> StencilObject:
        return compile_stencil(name, externals, self.gt4py_config)

    # -- helpers -------------------------------------------------------------------------
    def _raw_inputs(self, state: Dict[str, Any], props: PropertyDict) -> Dict[str, Any]:
        raw = {}
        for name, p in props.items():
            if name not in state:
                raise KeyError(f"{type(self).__name__}: input field {name!r} missing from the state")
            fld = state[name]
            if self.enable_checks and isinstance(fld, Field):
                if tuple(fld.grid_dims) != tuple(p["grid_dims"]):
                    raise ValueError(f"{type(self).__name__}: {name} has dims {fld.dims}, expected {p['grid_dims']}")
                # units are NOT enforced: the reference's own declarations disagree with each other
                # (f_tnd_cml_q is "g g^-1 s^-1" in nonlinear/microphysics.py:98 and "K s^-1" in
                # common/increment.py:66), so a strict check would reject the reference's own state
            raw[name] = fld.data if isinstance(fld, Field) else fld
        return raw

    def _outputs(self, out: Optional[Dict[str, Any]], props: PropertyDict) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        out = {} if out is None else out
        raw = {}
        for name, p in props.items():
            if name not in out:
                out[name] = allocate_field(self.computational_grid, name, p, self.gt4py_config)
            fld = out[name]
            raw[name] = fld.data if isinstance(fld, Field) else fld
        return out, raw


class DiagnosticComponent(_Component):
    input_grid_properties: PropertyDict
    diagnostic_grid_properties: PropertyDict

    def __call__(self, state: Dict[str, Any], *, out: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        raw_state = self._raw_inputs(state, self.input_grid_properties)
        out, raw_out = self._outputs(out, self.diagnostic_grid_properties)
        self.array_call(raw_state, raw_out)
        if "time" in state:
            out["time"] = state["time"]
        return out

    def array_call(self, state: Dict[str, Any], out: Dict[str, Any]) -> None:
        raise NotImplementedError


class ImplicitTendencyComponent(_Component):
    input_grid_properties: PropertyDict
    tendency_grid_properties: PropertyDict
    diagnostic_grid_properties: PropertyDict

    def __call__(self, state: Dict[str, Any], timestep: timedelta, *, out_tendencies: Optional[Dict[str, Any]] = None,
                 out_diagnostics: Optional[Dict[str, Any]] = None,
                 overwrite_tendencies: Optional[Dict[str, bool]] = None) -> Tuple[Dict[str, Any], Dict[str, Any]]:
        raw_state = self._raw_inputs(state, self.input_grid_properties)
        out_tendencies, raw_tends = self._outputs(out_tendencies, self.tendency_grid_properties)
        out_diagnostics, raw_diags = self._outputs(out_diagnostics, self.diagnostic_grid_properties)
        overwrite = overwrite_tendencies or {name: True for name in self.tendency_grid_properties}
        self.array_call(raw_state, timestep, raw_tends, raw_diags, overwrite)
        if "time" in state:
            out_tendencies["time"] = state["time"]
            out_diagnostics["time"] = state["time"]
        return out_tendencies, out_diagnostics

    def array_call(self, state: Dict[str, Any], timestep: timedelta, out_tendencies: Dict[str, Any],
                   out_diagnostics: Dict[str, Any], overwrite_tendencies: Dict[str, bool]) -> None:
        raise NotImplementedError
