"""The reference's components run on the reference's own stencil sources (TEST INFRASTRUCTURE).

`oracle/gtscript_exec.py` interprets the stencils; this module supplies what sits between a state dict and a
stencil call in the reference -- and takes that from the reference's files too, by reading (not importing: they
need sympl / ifs_physics_common) the component classes with `ast`:

  * the externals each component hands to `compile_stencil` -- its `__init__` is replayed statement by statement
    (`externals = {...}`, `externals.update(yoethf_params.dict())`, `externals.update({"ICALL": 0, ...})`, e.g.
    `nonlinear/microphysics.py:62-79`, `common/saturation.py:52-55`), the `*_params.dict()` contents being the
    fields the reference's parameter models declare (`iox.py:25-209`) looked up in the flat table `P`;
  * the keyword map of the stencil call in `array_call` (`in_ap=state["f_ap"]`, `out_tnd_q=out_tendencies["f_q"]`,
    `tmp_rfl=rfl`, `dt=...`, `domain=grids[I, J, K - 1/2].shape`; e.g. `nonlinear/microphysics.py:134-172`).

The functions below have the call signatures of `oracle/cloudsc2_numpy.py` so that tests can put the two side by
side: fields are `[nz+1, nx]` arrays, outputs are freshly zero-initialised storages like the reference's
`allocate_*` (the NL stencil relies on that for `out_fplsl[0]`).  Only usable where `/root/reference` exists.
"""
from __future__ import annotations

import ast
import os
from typing import Any, Dict, Optional, Tuple

import numpy as np

from . import gtscript_exec as gx

_PHYS = os.path.join(gx.REFERENCE_SRC, "cloudsc2_gt4py", "physics")

COMPONENTS = {
    "Saturation": "common/saturation.py",
    "StateIncrement": "common/increment.py",
    "PerturbedState": "common/increment.py",
    "Cloudsc2NL": "nonlinear/microphysics.py",
    "Cloudsc2TL": "tangent_linear/microphysics.py",
    "Cloudsc2AD": "adjoint/microphysics.py",
}

available = gx.available


def _class_node(cls: str) -> ast.ClassDef:
    with open(os.path.join(_PHYS, COMPONENTS[cls])) as f:
        tree = ast.parse(f.read())
    for n in tree.body:
        if isinstance(n, ast.ClassDef) and n.name == cls:
            return n
    raise KeyError(cls)


def _method(cls: ast.ClassDef, name: str) -> ast.FunctionDef:
    for n in cls.body:
        if isinstance(n, ast.FunctionDef) and n.name == name:
            return n
    raise KeyError(name)


_model_fields_cache: Optional[Dict[str, Dict[str, Any]]] = None


def model_fields() -> Dict[str, Dict[str, Any]]:
    """{model class name: {field: default or ...}} from the reference's `iox.py` parameter models."""
    global _model_fields_cache
    if _model_fields_cache is None:
        with open(os.path.join(gx.REFERENCE_SRC, "cloudsc2_gt4py", "iox.py")) as f:
            tree = ast.parse(f.read())
        out: Dict[str, Dict[str, Any]] = {}
        for n in tree.body:
            if isinstance(n, ast.ClassDef) and any(ast.unparse(b) == "BaseModel" for b in n.bases):
                out[n.name] = {
                    s.target.id: (ast.literal_eval(s.value) if s.value is not None else ...)
                    for s in n.body
                    if isinstance(s, ast.AnnAssign)
                }
        _model_fields_cache = out
    return _model_fields_cache


def _params_dict(arg_name: str, init: ast.FunctionDef, P: Dict[str, Any]) -> Dict[str, Any]:
    """`yoethf_params.dict()`: the annotated model class of that constructor argument, filled from P."""
    for a in init.args.args + init.args.kwonlyargs:
        if a.arg == arg_name:
            fields = model_fields()[ast.unparse(a.annotation)]
            out = {}
            for k, default in fields.items():
                if k in P:
                    out[k] = P[k]
                elif default is not ...:
                    out[k] = default
            return out
    raise KeyError(arg_name)


def component_externals(cls: str, P: Dict[str, Any], ctor: Dict[str, Any], nz: int) -> Tuple[str, Dict[str, Any]]:
    """Replays `<cls>.__init__` of the reference; returns (stencil name, externals)."""
    init = _method(_class_node(cls), "__init__")
    env = dict(ctor)
    env["nk"] = nz  # `nk = self.computational_grid.grids[I, J, K].shape[2]`, tangent_linear/microphysics.py:67
    externals: Dict[str, Any] = {}
    found = None
    for stmt in init.body:
        for call in [n for n in ast.walk(stmt) if isinstance(n, ast.Call)]:
            if ast.unparse(call.func) == "self.compile_stencil":
                name = ast.literal_eval(call.args[0])
                ext = externals
                for kw in call.keywords:
                    if kw.arg == "externals":
                        ext = eval(compile(ast.Expression(kw.value), "<ref>", "eval"), {}, env)
                found = (name, dict(ext))
        if isinstance(stmt, ast.Assign) and ast.unparse(stmt.targets[0]) == "externals":
            externals = eval(compile(ast.Expression(stmt.value), "<ref>", "eval"), {}, env)
        elif isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Call) and ast.unparse(stmt.value.func) == "externals.update":
            arg = stmt.value.args[0]
            if isinstance(arg, ast.Call) and isinstance(arg.func, ast.Attribute) and arg.func.attr == "dict":
                externals.update(_params_dict(arg.func.value.id, init, P))
            else:
                externals.update(eval(compile(ast.Expression(arg), "<ref>", "eval"), {}, env))
    assert found is not None, cls
    return found


def stencil_call_map(cls: str) -> Tuple[Dict[str, Tuple[str, Optional[str]]], bool]:
    """keyword -> (source, key) of the stencil call in `<cls>.array_call`; and whether the domain has nz+1 levels."""
    fn = _method(_class_node(cls), "array_call")
    calls = [
        n for n in ast.walk(fn)
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and ast.unparse(n.func.value) == "self"
        and any(kw.arg == "domain" for kw in n.keywords)
    ]
    assert len(calls) == 1, cls
    kwmap: Dict[str, Tuple[str, Optional[str]]] = {}
    half = False
    for kw in calls[0].keywords:
        v = kw.value
        if kw.arg == "domain":
            half = "K - 1 / 2" in ast.unparse(v)
        elif kw.arg in ("origin", "validate_args", "exec_info"):
            continue
        elif isinstance(v, ast.Subscript) and isinstance(v.value, ast.Name):
            kwmap[kw.arg] = (v.value.id, ast.literal_eval(v.slice))  # state["f_ap"], out_tendencies["f_q"]
        elif isinstance(v, ast.Name):
            kwmap[kw.arg] = ("tmp_ij", None)  # managed_temporary_storage IJ scratch
        elif ast.unparse(v) == "self.klevel":
            kwmap[kw.arg] = ("klevel", None)
        elif ast.unparse(v) == "self.f":
            kwmap[kw.arg] = ("factor", None)
        elif "total_seconds" in ast.unparse(v):
            kwmap[kw.arg] = ("dt", None)
        else:
            raise NotImplementedError(f"{cls}.array_call: {kw.arg}={ast.unparse(v)}")
    return kwmap, half


class Component:
    """`cls(grid, **ctor, <param models from P>)` of the reference, then `comp(state[, dt])`."""

    def __init__(self, cls: str, P: Dict[str, Any], nz: int, dtype, **ctor):
        self.cls, self.nz, self.dtype = cls, nz, np.dtype(dtype)
        name, ext = component_externals(cls, P, ctor, nz)
        self.stencil = gx.Stencil(name, ext, dtype)
        self.externals = ext
        self.kwmap, self.half = stencil_call_map(cls)
        self.factor = ctor.get("factor")

    def __call__(self, state: Dict[str, np.ndarray], dt: Optional[float] = None) -> Dict[str, Dict[str, np.ndarray]]:
        """Returns {"out" | "out_tendencies" | "out_diagnostics": {name: array}}; `state` arrays the stencil
        writes (the AD seeds) are modified in place, as in the reference."""
        nz = self.nz
        nx = next(v for k, v in state.items() if k != "f_eta" and isinstance(v, np.ndarray)).shape[1]
        outs: Dict[str, Dict[str, np.ndarray]] = {}
        kwargs: Dict[str, Any] = {}
        for arg, (src, key) in self.kwmap.items():
            if src == "state":
                kwargs[arg] = state[key]
            elif src in ("out", "out_tendencies", "out_diagnostics"):
                kwargs[arg] = outs.setdefault(src, {}).setdefault(key, np.zeros((nz + 1, nx), dtype=self.dtype))
            elif src == "tmp_ij":
                kwargs[arg] = np.zeros(nx, dtype=self.dtype)
            elif src == "klevel":
                kwargs[arg] = np.arange(0, nz + 1)  # tangent_linear/microphysics.py:68-71
            elif src == "factor":
                kwargs[arg] = self.dtype.type(self.factor)  # common/increment.py:46
            elif src == "dt":
                kwargs[arg] = self.dtype.type(dt)  # nonlinear/microphysics.py:167
        self.stencil(origin=(0, 0, 0), domain=(nx, 1, nz + 1 if self.half else nz), **kwargs)
        return outs


# ----------------------------------------------------------------------------------------------------------
# same signatures as oracle/cloudsc2_numpy.py
# ----------------------------------------------------------------------------------------------------------
def _nz(a: np.ndarray) -> int:
    return a.shape[0] - 1


def saturation(ap, t, P, kflag: Optional[int] = None):
    kflag = P.get("KFLAG", 1) if kflag is None else kflag
    comp = Component("Saturation", P, _nz(ap), ap.dtype, kflag=kflag, lphylin=bool(P["LPHYLIN"]))
    return comp({"f_ap": ap, "f_t": t})["out"]["f_qsat"]


def state_increment(state, f, ignore_supsat=False):
    a = state["f_ap"]
    comp = Component("StateIncrement", {}, _nz(a), a.dtype, factor=f, ignore_supsat=ignore_supsat)
    return comp(state)["out"]


def perturbed_state(state, f):
    a = state["f_ap"]
    comp = Component("PerturbedState", {}, _nz(a), a.dtype, factor=f)
    return comp(state)["out"]


def _microphysics(cls, s, dt, P):
    a = s["f_ap"]
    comp = Component(cls, P, _nz(a), a.dtype, lphylin=bool(P["LPHYLIN"]), ldrain1d=bool(P["LDRAIN1D"]))
    o = comp(s, dt)
    return o["out_tendencies"], o["out_diagnostics"]


def cloudsc2_nl(s, dt, P):
    return _microphysics("Cloudsc2NL", s, dt, P)


def cloudsc2_tl(s, dt, P):
    return _microphysics("Cloudsc2TL", s, dt, P)


def cloudsc2_ad(s, dt, P):
    """Literal reference behaviour (= the oracle's predicates="reference"); consumes the seeds in `s` in place."""
    return _microphysics("Cloudsc2AD", s, dt, P)
