"""Runs the reference's OWN gtscript stencil sources under NumPy (TEST INFRASTRUCTURE).

The reference's arithmetic lives in plain Python files written in GT4Py's `gtscript` DSL
(`/root/reference/src/cloudsc2_gt4py/physics/**/_stencils/*.py`); only the GT4Py *runtime* that would
compile them is missing from this image.  This module supplies the missing piece as an interpreter:

  * `load_reference()` imports the UNMODIFIED stencil files from `/root/reference/src` with two stub modules in
    place of their imports -- `gt4py.cartesian.gtscript` (`Field`, `K`, `IJ`, `function`) and
    `ifs_physics_common.stencil` (`stencil_collection(name)` / `function_collection(name)` register the python
    function under the reference's name, exactly what the real decorators do before handing it to GT4Py).
    Nothing is copied into this repository; the files are read where they lie.
  * `Stencil(name, externals, dtype)` interprets the registered function's AST with GT4Py cartesian semantics:
      - `with computation(FORWARD | BACKWARD | PARALLEL), interval(a, b)` blocks run one after another over the
        whole domain; FORWARD / BACKWARD are sequential in k, every statement is applied to a whole horizontal
        plane (vectorised over columns) before the next one -- the execution model of GT4Py's `numpy` backend,
        the reference default (`drivers/config.py:45`);
      - `F[0, 0, dk]` is a k-offset read, `F[0, 0]` an IJ field (one value per column that persists across
        levels), `F[0]` a K field; a bare field name means offset 0;
      - every local name of a stencil is a zero-initialised 3-D temporary (so it can be read at `[0, 0, -1]`
        in a later computation, e.g. `nonlinear/_stencils/cloudsc2.py:396`);
      - `from __externals__ import X` names are compile-time constants; an `if` whose test is a constant picks
        its branch, an `if` on field values is a per-point select (both branches evaluated, stores masked);
      - `gtscript.function`s are inlined: evaluated on the argument *values* with their own local scope, tuple
        returns supported.
    Arithmetic is NumPy in the field dtype (python-float literals and externals are weakly typed, NEP 50, so a
    float32 run stays in float32 throughout -- asserted on every operation).

It exists to pin `oracle/cloudsc2_numpy.py` (the hand restatement) and the committed fixtures to *outputs of the
reference's own source run here*: `tests/test_ref_exec.py` compares the two, `tests/golden/make_golden.py` writes
`tests/golden/ref_*.npz` from it.  `/root/reference` does not exist on the GPU box, so nothing GPU-side imports
this module; only the fixtures travel.  The product never imports it.
"""
from __future__ import annotations

import ast
import contextlib
import importlib
import inspect
import os
import sys
import textwrap
import types
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np

REFERENCE_SRC = os.environ.get("CS2_REFERENCE_SRC", "/root/reference/src")
_PKG = "cloudsc2_gt4py"
_FORMULATIONS = ("common", "nonlinear", "tangent_linear", "adjoint")

STENCILS: Dict[str, Callable] = {}
FUNCTIONS: Dict[str, Callable] = {}
_loaded = False


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, _PKG, "physics", "nonlinear", "_stencils"))


# ----------------------------------------------------------------------------------------------------------
# loading the reference's files with stubbed third-party imports
# ----------------------------------------------------------------------------------------------------------
class _Anything:
    """`gtscript.Field["float"]`, `gtscript.Field[gtscript.K, "float"]`: only the annotation *source* is used."""

    def __getitem__(self, item):
        return self


def _stub_modules() -> Dict[str, types.ModuleType]:
    mods: Dict[str, types.ModuleType] = {}

    def mod(name: str, **attrs) -> types.ModuleType:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        mods[name] = m
        return m

    gtscript = mod(
        "gt4py.cartesian.gtscript", Field=_Anything(), I="I", J="J", K="K", IJ="IJ", IK="IK", JK="JK", IJK="IJK",
        function=lambda f: f, FORWARD="FORWARD", BACKWARD="BACKWARD", PARALLEL="PARALLEL",
    )
    cartesian = mod("gt4py.cartesian", gtscript=gtscript)
    mod("gt4py", cartesian=cartesian)

    def stencil_collection(name):
        def deco(f):
            STENCILS[name] = f
            return f

        return deco

    def function_collection(name):
        def deco(f):
            FUNCTIONS[name] = f
            return f

        return deco

    st = mod("ifs_physics_common.stencil", stencil_collection=stencil_collection, function_collection=function_collection)
    ipc = mod("ifs_physics_common", stencil=st)
    ipc.__path__ = []  # a package, so that `ifs_physics_common.stencil` resolves through sys.modules
    # the reference's packages WITHOUT their __init__.py (those import the components, i.e. sympl & co.);
    # the `_stencils` sub-packages below them are imported for real, from the reference tree
    root = os.path.join(REFERENCE_SRC, _PKG)
    mod(_PKG).__path__ = [root]
    mod(_PKG + ".physics").__path__ = [os.path.join(root, "physics")]
    for f in _FORMULATIONS:
        mod(f"{_PKG}.physics.{f}").__path__ = [os.path.join(root, "physics", f)]
    mods["gt4py"].__path__ = []
    mods["gt4py.cartesian"].__path__ = []
    return mods


@contextlib.contextmanager
def _stubbed_imports():
    mods = _stub_modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        for k in [k for k in sys.modules if k == _PKG or k.startswith(_PKG + ".")]:
            sys.modules.pop(k, None)


def load_reference() -> None:
    """Imports every `_stencils` package of the reference (fills STENCILS / FUNCTIONS)."""
    global _loaded
    if _loaded:
        return
    if not available():
        raise FileNotFoundError(f"reference sources not found under {REFERENCE_SRC}")
    dont_write = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # /root/reference is read-only: no __pycache__ there
    try:
        with _stubbed_imports():
            for f in _FORMULATIONS:
                importlib.import_module(f"{_PKG}.physics.{f}._stencils")
    finally:
        sys.dont_write_bytecode = dont_write
    _loaded = True


def source_file(name: str) -> str:
    load_reference()
    return inspect.getsourcefile(STENCILS.get(name) or FUNCTIONS[name])


# ----------------------------------------------------------------------------------------------------------
# the interpreter
# ----------------------------------------------------------------------------------------------------------
def _fn_ast(fn: Callable) -> ast.FunctionDef:
    tree = ast.parse(textwrap.dedent(inspect.getsource(fn)))
    node = tree.body[0]
    assert isinstance(node, ast.FunctionDef), fn
    return node


_MATH = {
    "exp": np.exp, "log": np.log, "sqrt": np.sqrt, "tanh": np.tanh, "cosh": np.cosh, "sinh": np.sinh,
    "sin": np.sin, "cos": np.cos, "tan": np.tan, "abs": np.abs, "floor": np.floor, "ceil": np.ceil,
}

_BINOPS = {
    ast.Add: np.add, ast.Sub: np.subtract, ast.Mult: np.multiply, ast.Div: np.true_divide, ast.Pow: np.power,
}
_CMPOPS = {
    ast.Lt: np.less, ast.LtE: np.less_equal, ast.Gt: np.greater, ast.GtE: np.greater_equal, ast.Eq: np.equal,
    ast.NotEq: np.not_equal,
}


class _Return(Exception):
    def __init__(self, value):
        self.value = value


def _is_static(v) -> bool:
    return not isinstance(v, np.ndarray) or v.ndim == 0


class _Scope:
    """Name space of one inlined gtscript.function call: every local is a per-point value."""

    def __init__(self, values: Dict[str, Any]):
        self.values = values


class Stencil:
    """One registered reference stencil bound to externals and a float dtype (what `compile_stencil` returns)."""

    def __init__(self, name: str, externals: Dict[str, Any], dtype=np.float64, fn: Optional[Callable] = None):
        """`fn`: interpret this function instead of a registered reference stencil (the interpreter's own unit
        tests, tests/test_ref_exec.py, feed it toy stencils whose result is known in closed form)."""
        if fn is None:
            load_reference()
            fn = STENCILS[name]
        self.name = name
        self.fn = fn
        self.node = _fn_ast(self.fn)
        self.dtype = np.dtype(dtype)
        self.externals = {k: self._coerce_external(v) for k, v in externals.items()}
        self.kinds: Dict[str, str] = {}  # parameter name -> "IJK" | "IJ" | "K" | "scalar"
        for a in self.node.args.args:
            ann = ast.unparse(a.annotation)
            if "gtscript.IJ" in ann:
                self.kinds[a.arg] = "IJ"
            elif "gtscript.K" in ann:
                self.kinds[a.arg] = "K"
            else:
                assert "gtscript.Field" in ann, (a.arg, ann)
                self.kinds[a.arg] = "IJK"
        for a in self.node.args.kwonlyargs:
            self.kinds[a.arg] = "scalar"
        self._fn_cache: Dict[str, Tuple[ast.FunctionDef, Dict[str, Any]]] = {}
        self.statements_run = 0

    @staticmethod
    def _coerce_external(v):
        if isinstance(v, (bool, np.bool_)):
            return bool(v)
        if isinstance(v, (int, np.integer)):
            return int(v)
        return float(v)

    # ---------------------------------------------------------------- call
    def __call__(self, *, origin=(0, 0, 0), domain, validate_args=False, exec_info=None, **kwargs) -> None:
        """Fields are `[nk_storage, nx]` arrays (IJK), `[nx]` (IJ) or `[nk_storage]` (K); `domain = (nx, 1, nk)`."""
        assert tuple(origin) == (0, 0, 0)
        nx, ny, nk = domain
        assert ny == 1
        missing = set(self.kinds) - set(kwargs)
        extra = set(kwargs) - set(self.kinds)
        assert not missing and not extra, f"{self.name}: missing {sorted(missing)}, unexpected {sorted(extra)}"
        self.nx, self.nk = nx, nk
        self.args: Dict[str, Any] = {}
        for name, kind in self.kinds.items():
            v = kwargs[name]
            if kind == "scalar":
                self.args[name] = self.dtype.type(v)
                continue
            assert isinstance(v, np.ndarray), name
            if kind == "IJK":
                assert v.ndim == 2 and v.shape[0] >= nk and v.shape[1] == nx and v.dtype == self.dtype, (name, v.shape, v.dtype)
            elif kind == "IJ":
                assert v.shape == (nx,) and v.dtype == self.dtype, (name, v.shape, v.dtype)
            else:
                assert v.ndim == 1 and v.shape[0] >= nk, (name, v.shape)
            self.args[name] = v
        self.temps: Dict[str, np.ndarray] = {}
        with np.errstate(all="ignore"):
            for stmt in self.node.body:
                if isinstance(stmt, ast.With):
                    self._computation(stmt)
                else:
                    assert isinstance(stmt, (ast.ImportFrom, ast.Expr)), ast.dump(stmt)

    # ---------------------------------------------------------------- computations and intervals
    def _interval(self, call: ast.Call) -> range:
        assert isinstance(call, ast.Call) and call.func.id == "interval"
        a = [ast.literal_eval(x) for x in call.args]
        if a == [Ellipsis]:
            return range(0, self.nk)
        lo, hi = a
        lo = self.nk + lo if lo < 0 else lo
        hi = self.nk if hi is None else (self.nk + hi if hi < 0 else hi)
        return range(lo, max(lo, hi))

    def _computation(self, node: ast.With) -> None:
        comp = node.items[0].context_expr
        assert isinstance(comp, ast.Call) and comp.func.id == "computation", ast.dump(comp)
        order = comp.args[0].id
        assert order in ("FORWARD", "BACKWARD", "PARALLEL")
        if len(node.items) == 2:
            blocks = [(self._interval(node.items[1].context_expr), node.body)]
        else:
            blocks = []
            for sub in node.body:
                assert isinstance(sub, ast.With) and len(sub.items) == 1, ast.dump(sub)
                blocks.append((self._interval(sub.items[0].context_expr), sub.body))
        covered: Dict[int, List[ast.stmt]] = {}
        for rng, body in blocks:
            for k in rng:
                assert k not in covered, f"{self.name}: overlapping intervals at k={k}"
                covered[k] = body
        if order == "PARALLEL":
            self._check_parallel_is_pointwise(blocks)
        ks = sorted(covered, reverse=(order == "BACKWARD"))
        for k in ks:
            self.k = k
            self._block(covered[k], None, None)

    def _check_parallel_is_pointwise(self, blocks) -> None:
        """A PARALLEL block is run level by level here; that equals GT4Py's statement-by-statement order only
        if nothing written in the block is read at a vertical offset in it."""
        for _, body in blocks:
            written = set()
            for n in ast.walk(ast.Module(body=body, type_ignores=[])):
                if isinstance(n, (ast.Assign, ast.AugAssign)):
                    for t in n.targets if isinstance(n, ast.Assign) else [n.target]:
                        for e in t.elts if isinstance(t, ast.Tuple) else [t]:
                            written.add(e.id if isinstance(e, ast.Name) else e.value.id)
            for n in ast.walk(ast.Module(body=body, type_ignores=[])):
                if isinstance(n, ast.Subscript) and isinstance(n.ctx, ast.Load) and n.value.id in written:
                    off = ast.literal_eval(n.slice)
                    assert not any(off if isinstance(off, tuple) else (off,)), f"offset read of {n.value.id} in PARALLEL"

    # ---------------------------------------------------------------- statements
    def _block(self, body: List[ast.stmt], mask: Optional[np.ndarray], scope: Optional[_Scope]) -> None:
        for stmt in body:
            self.statements_run += 1
            if isinstance(stmt, ast.Assign):
                assert len(stmt.targets) == 1
                self._assign(stmt.targets[0], self._eval(stmt.value, scope), mask, scope)
            elif isinstance(stmt, ast.AugAssign):
                cur = self._eval(self._as_load(stmt.target), scope)
                val = self._binop(type(stmt.op), cur, self._eval(stmt.value, scope))
                self._assign(stmt.target, val, mask, scope)
            elif isinstance(stmt, ast.If):
                self._if(stmt, mask, scope)
            elif isinstance(stmt, ast.Return):
                assert scope is not None and mask is None, "return under a per-point condition"
                raise _Return(self._eval(stmt.value, scope))
            elif isinstance(stmt, (ast.ImportFrom, ast.Pass)):
                continue
            elif isinstance(stmt, ast.Expr) and isinstance(stmt.value, ast.Constant):
                continue  # docstring
            else:
                raise NotImplementedError(ast.dump(stmt))

    @staticmethod
    def _as_load(target: ast.expr) -> ast.expr:
        if isinstance(target, ast.Name):
            return ast.Name(id=target.id, ctx=ast.Load())
        return ast.Subscript(value=target.value, slice=target.slice, ctx=ast.Load())

    def _if(self, node: ast.If, mask, scope) -> None:
        cond = self._eval(node.test, scope)
        if _is_static(cond):
            self._block(node.body if bool(cond) else node.orelse, mask, scope)
            return
        assert cond.dtype == np.bool_ and cond.shape == (self.nx,), (ast.unparse(node.test), cond.dtype, cond.shape)
        # GT4Py evaluates the test once, before either branch runs
        m_then = cond if mask is None else (mask & cond)
        m_else = ~cond if mask is None else (mask & ~cond)
        self._block(node.body, m_then, scope)
        if node.orelse:
            self._block(node.orelse, m_else, scope)

    # ---------------------------------------------------------------- stores
    def _typed(self, value, what: str):
        """Values stored in the float dtype (or bool for logical temporaries); catches silent promotion."""
        if isinstance(value, np.ndarray) and value.ndim > 0:
            if value.dtype == np.bool_ or value.dtype == self.dtype:
                return value
            raise TypeError(f"{self.name}: {what} evaluated in {value.dtype}, not {self.dtype}")
        if isinstance(value, (bool, np.bool_)):
            return np.bool_(value)
        if isinstance(value, np.floating) and value.dtype != self.dtype:
            raise TypeError(f"{self.name}: {what} evaluated in {value.dtype}, not {self.dtype}")
        return self.dtype.type(value)

    def _assign(self, target: ast.expr, value, mask, scope) -> None:
        if isinstance(target, ast.Tuple):
            assert isinstance(value, tuple) and len(value) == len(target.elts)
            for t, v in zip(target.elts, value):
                self._assign(t, v, mask, scope)
            return
        if isinstance(target, ast.Name):
            name, off = target.id, None
        else:
            assert isinstance(target, ast.Subscript) and isinstance(target.value, ast.Name)
            name, off = target.value.id, ast.literal_eval(target.slice)
            assert not any(off if isinstance(off, tuple) else (off,)), f"store at an offset: {ast.unparse(target)}"
        value = self._typed(value, name)
        if scope is not None:  # local of an inlined function
            assert off is None
            if mask is None:
                scope.values[name] = value
            else:
                old = scope.values.get(name)
                if old is None:
                    old = np.zeros((), dtype=value.dtype)
                scope.values[name] = np.where(mask, value, old)
            return
        kind = self.kinds.get(name)
        if kind is None:  # 3-D temporary of the stencil, zero-initialised
            assert name not in self.externals, f"store to external {name}"
            arr = self.temps.get(name)
            if arr is None:
                arr = self.temps[name] = np.zeros((self.nk, self.nx), dtype=value.dtype)
            dst = arr[self.k]
        elif kind == "IJK":
            assert off in ((0, 0, 0),), f"{name}: 3-D field stored without [0, 0, 0]"
            dst = self.args[name][self.k]
        elif kind == "IJ":
            assert off == (0, 0)
            dst = self.args[name]
        else:
            raise NotImplementedError(f"store to {kind} argument {name}")
        assert dst.dtype == value.dtype, f"{self.name}: {name} is {dst.dtype}, value is {value.dtype}"
        if mask is None:
            dst[...] = value
        else:
            np.copyto(dst, value, where=mask)

    # ---------------------------------------------------------------- expressions
    def _eval(self, node: ast.expr, scope: Optional[_Scope]):
        if isinstance(node, ast.Constant):
            assert isinstance(node.value, (bool, int, float)), node.value
            return node.value
        if isinstance(node, ast.Name):
            return self._load(node.id, None, scope)
        if isinstance(node, ast.Subscript):
            assert isinstance(node.value, ast.Name)
            return self._load(node.value.id, ast.literal_eval(node.slice), scope)
        if isinstance(node, ast.BinOp):
            return self._binop(type(node.op), self._eval(node.left, scope), self._eval(node.right, scope))
        if isinstance(node, ast.UnaryOp):
            v = self._eval(node.operand, scope)
            if isinstance(node.op, ast.USub):
                return -v
            if isinstance(node.op, ast.UAdd):
                return v
            assert isinstance(node.op, ast.Not)
            return (not v) if isinstance(v, bool) else np.logical_not(v)
        if isinstance(node, ast.BoolOp):
            vals = [self._eval(v, scope) for v in node.values]
            fn = np.logical_and if isinstance(node.op, ast.And) else np.logical_or
            out = vals[0]
            for v in vals[1:]:
                if isinstance(out, bool) and isinstance(v, bool):
                    out = (out and v) if isinstance(node.op, ast.And) else (out or v)
                else:
                    out = fn(out, v)
            return out
        if isinstance(node, ast.Compare):
            assert len(node.ops) == 1
            a, b = self._eval(node.left, scope), self._eval(node.comparators[0], scope)
            return _CMPOPS[type(node.ops[0])](a, b)
        if isinstance(node, ast.Call):
            return self._call(node, scope)
        if isinstance(node, ast.Tuple):
            return tuple(self._eval(e, scope) for e in node.elts)
        if isinstance(node, ast.IfExp):
            c = self._eval(node.test, scope)
            a, b = self._eval(node.body, scope), self._eval(node.orelse, scope)
            return (a if bool(c) else b) if _is_static(c) else np.where(c, a, b)
        raise NotImplementedError(ast.dump(node))

    def _binop(self, op, a, b):
        if all(isinstance(x, (bool, int, float)) for x in (a, b)):  # compile-time arithmetic on literals/externals
            if op is ast.Add:
                return a + b
            if op is ast.Sub:
                return a - b
            if op is ast.Mult:
                return a * b
            if op is ast.Div:
                return a / b
            return a**b
        r = _BINOPS[op](a, b)
        if isinstance(r, (np.ndarray, np.floating)) and r.dtype.kind == "f" and r.dtype != self.dtype:
            raise TypeError(f"{self.name}: {op.__name__} promoted to {r.dtype}")
        return r

    def _load(self, name: str, off, scope: Optional[_Scope]):
        if scope is not None and name in scope.values:
            assert off is None
            return scope.values[name]
        kind = None if scope is not None else self.kinds.get(name)
        if kind == "IJK":
            dk = 0 if off is None else self._dk(off, 3)
            return self.args[name][self.k + dk]
        if kind == "IJ":
            assert off in (None, (0, 0))
            return self.args[name]
        if kind == "K":
            dk = 0 if off is None else self._dk(off, 1)
            return self.args[name][self.k + dk]
        if kind == "scalar":
            return self.args[name]
        if scope is None and name in self.temps:
            dk = 0 if off is None else self._dk(off, 3)
            k = self.k + dk
            assert 0 <= k < self.nk, f"{name}[{off}] read outside the domain at k={self.k}"
            return self.temps[name][k]
        if name in self.externals:
            assert off is None
            return self.externals[name]
        raise NameError(f"{self.name}: `{name}` read before any assignment (k={self.k})")

    @staticmethod
    def _dk(off, rank: int) -> int:
        if rank == 1:
            return off if isinstance(off, int) else off[0]
        assert isinstance(off, tuple) and len(off) == 3 and off[0] == 0 and off[1] == 0, off
        return off[2]

    def _call(self, node: ast.Call, scope: Optional[_Scope]):
        assert isinstance(node.func, ast.Name) and not node.keywords, ast.dump(node)
        fname = node.func.id
        args = [self._eval(a, scope) for a in node.args]
        if fname in ("min", "max"):
            assert len(args) == 2
            if all(isinstance(x, (int, float)) for x in args):
                return min(*args) if fname == "min" else max(*args)
            r = (np.minimum if fname == "min" else np.maximum)(*args)
            assert r.dtype == self.dtype, (fname, r.dtype)
            return r
        if fname in _MATH:
            assert len(args) == 1
            a = args[0]
            if isinstance(a, (int, float)):
                a = self.dtype.type(a)
            r = _MATH[fname](a)
            assert r.dtype == self.dtype, (fname, r.dtype)
            return r
        return self._inline(fname, args)

    def _inline(self, fname: str, args: List[Any]):
        entry = self._fn_cache.get(fname)
        if entry is None:
            fn = self.fn.__globals__.get(fname)
            if fn is None:  # functions find each other through the module they were defined in
                fn = FUNCTIONS[fname]
            entry = self._fn_cache[fname] = (_fn_ast(fn), fn.__globals__)
        fnode, fglobals = entry
        params = [a.arg for a in fnode.args.args]
        assert len(params) == len(args), fname
        outer_fn = self.fn
        self.fn = types.SimpleNamespace(__globals__=fglobals)  # nested calls resolve in the callee's module
        try:
            self._block(fnode.body, None, _Scope(dict(zip(params, args))))
        except _Return as r:
            return r.value
        finally:
            self.fn = outer_fn
        raise RuntimeError(f"{fname} did not return")
