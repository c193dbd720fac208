"""Lowering of real Solverz models into this backend (the seam of ``code_printer/make_module.py:28-37``).

``from_solverz(eqs, y0)`` takes what ``Solverz.Model.create_instance()`` returns -- an ``AE`` / ``DAE`` /
``FDAE`` (``Solverz/equation/equations.py``) and its ``Vars`` -- and rebuilds it as this package's
``Equations`` + ``Vars``, so that

    from Solverz import Model, Var, Param, Eqn, ...          # the user's model, untouched
    spf, y0 = m.create_instance()
    module_printer(spf, y0, 'name', lang='cuda').render()    # solverz_b200.module_printer
    made_numerical(spf, y0, sparse=True)                     # solverz_b200.made_numerical

work on the SAME Python model object both back-ends accept.  What is read from the reference objects:

* variables and their order: ``eqs.var_address`` (``utilities/address.py``) and the values in ``y0``;
* parameters: ``eqs.PARAM`` (``equation/param.py:12-212``: ``Param``, ``IdxParam``, ``TimeSeriesParam``, alias
  parameters of ``AliasVar``); the auxiliary ``<A>_data/_indices/_indptr/_shape0`` entries of sparse matrices
  (``model/basic.py:56-60``) are dropped, the matrix itself is lowered to CSR row walkers;
* equations in declaration order (``eqs.EQNs``; a DAE's ``Ode`` first, ``equations.py:672``): the sympy tree
  ``Eqn.RHS`` over ``iVar / Para / IdxVar / IdxPara / iAliasVar`` leaves (``sym_algebra/symbols.py:101-246``) and the
  function classes of ``sym_algebra/functions.py``; ``LoopEqn`` / ``LoopOde`` bodies over
  ``sympy.IndexedBase / Idx / Sum`` with their ``var_map`` (``equation/eqn.py:890-1147``), subset ``Set`` gathers
  included.

The Jacobian is NOT taken from the reference's ``FormJac`` blocks: this backend differentiates the lowered row
templates itself (``solverz_b200/jacobian.py``) and produces the same CSC pattern and values (parity tests:
``tests/test_cpu_adapter.py`` against the reference's own ``F_`` / ``J_``).

``Mat_Mul`` chains (``functions.py:40-124``): ``Mat_Mul(A, x)``, ``Mat_Mul(-A, x)``, ``Mat_Mul(c*A, x)``, nested and
multi-argument products of FROZEN matrices (``Mat_Mul(A, Mat_Mul(B, x))``, ``Mat_Mul(A, B, x)``: the product matrix is
formed once on the host -- 2-D parameters are immutable in the reference too), and ``Diag(v)`` factors on either
side of one matrix (``Mat_Mul(Diag(u), A, Diag(v), x) = u * (A @ (v * x))``).  A matrix applied to a vector that
itself depends NON-linearly on another matrix product is rejected with ``NotImplementedError``.  ``Ones(n)``
broadcasts (``functions.py:726-749``) and becomes the scalar 1 of the row template.

Triggerable parameters (``param.py:16-59``: ``Param(..., triggerable=True, trigger_var=[...], trigger_fun=f)``): the
reference's generated ``F_`` / ``J_`` overwrite the parameter with ``f(trigger variables)`` at every call
(``code_printer/python/utilities.py:118-129``), and the Jacobian treats it as a parameter (no derivative through
``f``).  Here the parameter is replaced, inside the expressions, by ``frozen(f(...))`` -- a function whose value is its
argument and whose derivative is zero -- when ``f`` can be written symbolically: a numpy ufunc of the table below, or
any callable that accepts symbols (``lambda x: 2 * x + 1``).  Opaque host callbacks are rejected with a message.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import scipy.sparse as sps
import sympy as sp

from . import sym as S
from .equations import AE, DAE, FDAE, Eqn, Equations, LoopEqn, LoopOde, Ode
from .variables import Vars

__all__ = ["from_solverz", "is_solverz_object", "frozen"]


class frozen(sp.Function):
    """``frozen(e)`` has the value of ``e`` and no derivative: how a triggerable parameter ``p = f(vars)`` of the
    reference enters F (its current value) and J (a constant)."""
    nargs = 1
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)

    def _eval_derivative(self, s):
        return sp.Integer(0)

    def _cudacode(self, printer):
        return f"({printer._print(self.args[0])})"


# numpy ufuncs a trigger_fun may be (called on symbols they raise); anything else is tried on the symbols directly
_TRIGGER_UFUNCS = {"sin": sp.sin, "cos": sp.cos, "tan": sp.tan, "exp": sp.exp, "log": sp.log, "sqrt": sp.sqrt,
                   "absolute": sp.Abs, "fabs": sp.Abs, "sign": S.Sign, "tanh": sp.tanh, "sinh": sp.sinh,
                   "cosh": sp.cosh, "arctan": sp.atan, "arctan2": sp.atan2, "square": lambda x: x ** 2,
                   "negative": lambda x: -x, "reciprocal": lambda x: 1 / x}


def _trigger_expression(name: str, p, variables: Dict[str, S.Var]):
    """The symbolic form of ``p.trigger_fun(*trigger variables)`` over this backend's variable symbols."""
    tv = list(p.trigger_var or [])
    fun = p.trigger_fun
    if not tv or fun is None:
        raise ValueError(f"parameter {name}: triggerable=True needs trigger_var and trigger_fun")
    missing = [v for v in tv if v not in variables]
    if missing:
        raise NotImplementedError(f"parameter {name}: trigger_var {missing} are not variables of the model "
                                  f"(parameters as trigger inputs are not lowered)")
    args = [variables[v].sym for v in tv]
    lens = {variables[v].size for v in tv if variables[v].size != 1}
    size = int(np.size(p.v)) if p.v is not None else None
    if size is not None and size != 1 and lens and lens != {size}:
        raise NotImplementedError(f"parameter {name}: trigger variables of length {sorted(lens)} for a parameter of "
                                  f"length {size}; only elementwise trigger functions are lowered")
    uf = getattr(fun, "__name__", "") if isinstance(fun, np.ufunc) else None
    if uf is not None:
        if uf not in _TRIGGER_UFUNCS:
            raise NotImplementedError(f"parameter {name}: numpy ufunc {uf} has no symbolic counterpart here")
        return _TRIGGER_UFUNCS[uf](*args)
    try:
        e = fun(*args)
    except Exception as ex:
        raise NotImplementedError(f"parameter {name}: trigger_fun {getattr(fun, '__name__', fun)!r} is an opaque host "
                                  f"callback ({type(ex).__name__} when called on symbols); the device cannot call "
                                  f"it.  Write it with operators / solverz_b200 functions, or update mdl.p[{name!r}] "
                                  f"between solves instead") from None
    try:
        return sp.sympify(e)
    except Exception:
        raise NotImplementedError(f"parameter {name}: trigger_fun returned {type(e).__name__}, not an expression")

_MATRIX_HELPER_SUFFIXES = ("_data", "_indices", "_indptr", "_shape0")

# reference function class name -> constructor in this backend's vocabulary
_FUNCS = {
    "sin": sp.sin, "cos": sp.cos, "exp": sp.exp, "ln": sp.log, "Abs": sp.Abs, "atan2": sp.atan2,
    "Sign": S.Sign, "heaviside": S.heaviside, "Not": S.Not, "switch": S.switch, "Saturation": S.Saturation,
    "AntiWindUp": S.AntiWindUp, "Min": S.Min, "In": S.In, "GreaterThan": S.GreaterThan, "LessThan": S.LessThan,
    "And": S.And, "Or": S.Or,
    # sympy's own classes may appear as well (sympify of user expressions)
    "log": sp.log, "tan": sp.tan, "sqrt": sp.sqrt, "sign": S.Sign,
}


def is_solverz_object(obj) -> bool:
    """True for an equation container of the reference (``Solverz.equation.equations``)."""
    mod = type(obj).__module__ or ""
    return mod.startswith("Solverz.") and hasattr(obj, "EQNs") and hasattr(obj, "PARAM")


class _Translator:
    def __init__(self, variables: Dict[str, S.Var], params: Dict[str, S.Param], eq_name: str, triggered=None):
        self.vars = variables
        self.params = params
        self.eq = eq_name
        self.triggered = triggered or {}        # name -> symbolic f(trigger variables)
        self._counter = 0

    # ---- leaves -------------------------------------------------------------------------------
    def _base(self, name: str):
        if name in self.vars:
            return self.vars[name]
        if name in self.params:
            return self.params[name]
        raise ValueError(f"{self.eq}: symbol {name!r} is neither a variable nor a parameter of the model")

    @staticmethod
    def _int(v, what):
        if v is None:
            return None
        if isinstance(v, (int, np.integer)):
            return int(v)
        if isinstance(v, sp.Integer):
            return int(v)
        raise NotImplementedError(f"index {what}={v!r}: only integer slice bounds are supported "
                                  f"(the reference's numba printer rejects symbolic indices as well, "
                                  f"code_printer/python/utilities.py:257-261)")

    def _plain_index(self, index):
        if isinstance(index, (int, np.integer, sp.Integer)):
            return int(index)
        if isinstance(index, slice):
            return slice(self._int(index.start, "start"), self._int(index.stop, "stop"), self._int(index.step, "step"))
        raise NotImplementedError(f"{self.eq}: index {index!r} ({type(index).__name__}); use int or slice "
                                  f"(list / idx fancy indices are rejected by the reference's module printer too)")

    # ---- ordinary equations --------------------------------------------------------------------
    def expr(self, e):
        cname = type(e).__name__
        if isinstance(e, (sp.Number, sp.NumberSymbol)):
            return e
        if cname in ("iVar", "Para", "iAliasVar", "idx"):
            if e.name in self.triggered:
                return frozen(self.triggered[e.name])
            b = self._base(e.name)
            if isinstance(b, S.Param) and b.dim == 2:
                raise NotImplementedError(f"{self.eq}: the 2-D parameter {e.name} may only appear as a factor of "
                                          f"Mat_Mul")
            return b.sym
        if cname in ("IdxVar", "IdxPara", "IdxAliasVar", "Idxidx"):
            if e.name0 in self.triggered:
                raise NotImplementedError(f"{self.eq}: the triggerable parameter {e.name0} is used with an index; only "
                                          f"whole-parameter uses are lowered")
            return self._base(e.name0)[self._plain_index(e.index)]
        if isinstance(e, sp.Symbol):
            if e.name == "t":
                return S.t
            raise ValueError(f"{self.eq}: unknown symbol {e!r} ({cname})")
        if cname == "Mat_Mul":
            return self.matmul(list(e.args))
        if cname == "Ones":
            return sp.Integer(1)                       # broadcast helper: 2*Ones(n) is the scalar 2 per row
        if cname in ("Diag", "SpDiag", "transpose"):
            raise NotImplementedError(f"{self.eq}: {cname} outside a Mat_Mul chain")
        if cname in _FUNCS and isinstance(e, sp.Function):
            return _FUNCS[cname](*[self.expr(a) for a in e.args])
        if isinstance(e, (sp.Add, sp.Mul, sp.Pow)):
            return e.func(*[self.expr(a) for a in e.args])
        if isinstance(e, sp.Function):
            if hasattr(e, "_cudacode"):                # user function with a CUDA printer hook
                return e.func(*[self.expr(a) for a in e.args])
            raise NotImplementedError(f"{self.eq}: function {cname} has no CUDA counterpart (custom functions need "
                                      f"_cudacode / _cuda_device_code, the analogue of _numpycode)")
        raise NotImplementedError(f"{self.eq}: cannot lower {e!r} ({cname})")

    # ---- Mat_Mul chains --------------------------------------------------------------------------
    def _matrix_factor(self, a) -> Optional[Tuple[float, sps.csr_matrix, str]]:
        """(scalar, matrix, label) when ``a`` is c * (frozen 2-D parameter), else None."""
        c = 1.0
        if isinstance(a, sp.Mul):
            nums = [x for x in a.args if isinstance(x, sp.Number)]
            rest = [x for x in a.args if not isinstance(x, sp.Number)]
            if len(rest) != 1:
                return None
            for x in nums:
                c *= float(x)
            a = rest[0]
        if type(a).__name__ == "Para":              # the symbol's own .dim is not reliable after FormJac: ask the model
            p = self.params.get(a.name)
            if p is not None and p.dim == 2:
                return c, p.value, a.name
        return None

    def _derived_matrix(self, M: sps.spmatrix, label: str) -> S.Param:
        name = f"_szad_{label}"
        if name in self.params:
            old = self.params[name].value
            if old.shape == M.shape and (abs(old - M)).nnz == 0:
                return self.params[name]
            self._counter += 1
            name = f"{name}_{self._counter}"
        p = S.Param(name, sps.csr_matrix(M), dim=2, sparse=True)
        self.params[name] = p
        return p

    def matmul(self, args: List):
        """args[0] @ args[1] @ ... @ args[-1], the last one a vector expression."""
        vec = args[-1]
        # state: either a plain vector expression `e`, or a pending product  c * (M @ e)  whose matrix can still
        # absorb further frozen matrices from the left
        if type(vec).__name__ == "Mat_Mul":
            pend = self._pending(list(vec.args))
        else:
            pend = (None, 1.0, "", self.expr(vec))
        for a in reversed(args[:-1]):
            pend = self._apply_left(a, pend)
        return self._close(pend)

    def _pending(self, args: List):
        vec = args[-1]
        pend = self._pending(list(vec.args)) if type(vec).__name__ == "Mat_Mul" else (None, 1.0, "", self.expr(vec))
        for a in reversed(args[:-1]):
            pend = self._apply_left(a, pend)
        return pend

    def _close(self, pend):
        M, c, label, e = pend
        if M is None:
            return c * e if c != 1.0 else e
        p = self._derived_matrix(M, label) if "@" in label or c != 1.0 and False else None
        if p is None:
            p = self.params.get(label)
            if p is None or p.dim != 2:
                p = self._derived_matrix(M, label)
        out = S.Mat_Mul(p, e)
        return c * out if c != 1.0 else out

    def _apply_left(self, a, pend):
        M, c, label, e = pend
        cname = type(a).__name__
        if cname in ("Diag", "SpDiag"):
            d = self.expr(a.args[0])
            if M is None:
                return (None, c, "", d * e)
            return (None, 1.0, "", d * self._close(pend))       # u * (M @ e): the walker symbol times a row value
        mf = self._matrix_factor(a)
        if mf is None:
            raise NotImplementedError(f"{self.eq}: Mat_Mul factor {a!r} is neither a frozen 2-D parameter (times a "
                                      f"number) nor Diag(...)")
        ca, A, la = mf
        if M is None:
            if any(S.sum_info(s) is not None for s in sp.sympify(e).free_symbols):
                raise NotImplementedError(
                    f"{self.eq}: {la} is applied to a vector that depends on another matrix product through a "
                    f"non-linear or Diag-scaled expression; the row-template IR has no nested walks "
                    f"(products of frozen matrices alone are folded on the host)")
            return (A, c * ca, la, e)
        return (sps.csr_matrix(A @ M), c * ca, f"{la}@{label}".replace("@", "_x_") + "@", e)

    # ---- LoopEqn bodies --------------------------------------------------------------------------
    def loop_body(self, body, outer_ref, outer_mine: S.Idx, inner: Dict[str, S.Idx]):
        e = body
        if isinstance(e, (sp.Number, sp.NumberSymbol)):
            return e
        if isinstance(e, sp.Sum):
            f, (j, lo, hi) = e.args[0], e.args[1]
            if len(e.args) != 2 or int(lo) != 0:
                raise NotImplementedError(f"{self.eq}: Sum must run over one index from 0")
            jm = S.Idx(str(j), int(hi) + 1)
            inner2 = dict(inner)
            inner2[str(j)] = jm
            return S.Sum(self.loop_body(f, outer_ref, outer_mine, inner2), jm)
        if isinstance(e, sp.Indexed):
            base = self._base(str(e.base))
            keys = [self._loop_index(k, outer_ref, outer_mine, inner) for k in e.indices]
            return base[keys[0]] if len(keys) == 1 else base[tuple(keys)]
        if isinstance(e, sp.Idx):
            raise NotImplementedError(f"{self.eq}: a bare index {e} as a value")
        cname = type(e).__name__
        if cname in ("iVar", "Para", "iAliasVar"):
            raise NotImplementedError(f"{self.eq}: whole-vector symbol {e} inside a LoopEqn body")
        if cname in _FUNCS and isinstance(e, sp.Function):
            return _FUNCS[cname](*[self.loop_body(a, outer_ref, outer_mine, inner) for a in e.args])
        if isinstance(e, (sp.Add, sp.Mul, sp.Pow)):
            return e.func(*[self.loop_body(a, outer_ref, outer_mine, inner) for a in e.args])
        if isinstance(e, sp.Symbol) and e.name == "t":
            return S.t
        raise NotImplementedError(f"{self.eq}: cannot lower {e!r} ({cname}) inside a LoopEqn body")

    def _loop_index(self, k, outer_ref, outer_mine, inner):
        if isinstance(k, (int, np.integer, sp.Integer)):
            return int(k)
        if isinstance(k, sp.Idx):
            if str(k) == str(outer_ref):
                return outer_mine
            if str(k) in inner:
                return inner[str(k)]
            raise ValueError(f"{self.eq}: index {k} is neither the outer index nor bound by a Sum")
        if isinstance(k, sp.Indexed):                  # Set / int-parameter gather: PVPQ[i]
            par = self._base(str(k.base))
            if not (isinstance(par, S.Param) and par.is_int) or len(k.indices) != 1:
                raise NotImplementedError(f"{self.eq}: indirect index {k} must be a 1-D integer parameter")
            return par[self._loop_index(k.indices[0], outer_ref, outer_mine, inner)]
        raise NotImplementedError(f"{self.eq}: index expression {k!r}")


def _param_from_reference(name: str, p, per_scenario: bool = False) -> S.Param:
    cname = type(p).__name__
    dim = int(getattr(p, "dim", 1))
    if cname == "TimeSeriesParam":
        if dim != 1:
            raise NotImplementedError(f"parameter {name}: 2-D TimeSeriesParam")
        return S.TimeSeriesParam(name, np.asarray(p.v_series, dtype=float), np.asarray(p.time_series, dtype=float),
                                 value=np.asarray(p.v, dtype=float), index=p.index, per_scenario=per_scenario)
    if getattr(p, "is_alias", False):
        base, _, k = name.rpartition("_tag_")
        return S.AliasVar(base, step=int(k) + 1, value=None if p.v is None else np.asarray(p.v, dtype=float))
    is_int = cname == "IdxParam" or getattr(p, "dtype", float) is int
    if dim == 2:
        v = p.v
        M = sps.csr_matrix(v) if sps.issparse(v) else sps.csr_matrix(np.asarray(v, dtype=float))
        return S.Param(name, M, dim=2, sparse=True)
    if per_scenario and is_int:
        raise ValueError(f"parameter {name}: integer (index) parameters are shared by all scenarios")
    return S.Param(name, np.asarray(p.v), dtype=int if is_int else float, per_scenario=per_scenario and not is_int)


def from_solverz(eqs, y0, per_scenario=()) -> Tuple[Equations, Vars]:
    """Lower a Solverz ``AE`` / ``DAE`` / ``FDAE`` and its ``Vars`` (module docstring).

    ``per_scenario``: names of the 1-D parameters (``Param`` values, or the knots of a ``TimeSeriesParam``) that
    differ between the scenarios of a batch -- the one declaration the reference has no word for (there a user
    mutates ``mdl.p[name]`` between solves); ``mdl.p[name]`` then takes ``[B, len]`` values.  Everything else is
    shared by the batch."""
    if not is_solverz_object(eqs):
        raise TypeError(f"from_solverz expects a Solverz AE / DAE / FDAE, got {type(eqs).__name__}")
    per_scenario = set(per_scenario or ())
    unknown = per_scenario - set(eqs.PARAM)
    if unknown:
        raise ValueError(f"per_scenario names that are not parameters of the model: {sorted(unknown)}")
    kind = type(eqs).__name__
    cls = {"AE": AE, "DAE": DAE, "FDAE": FDAE}.get(kind)
    if cls is None:
        raise TypeError(f"unsupported equation container {kind}")
    # variables, in address order
    addr = eqs.var_address
    names = list(addr.v.keys()) if hasattr(addr, "v") else list(addr.object_list)
    names.sort(key=lambda nm: int(np.min(addr.v[nm])) if len(addr.v[nm]) else 0)
    variables: Dict[str, S.Var] = {}
    for nm in names:
        variables[nm] = S.Var(nm, np.asarray(y0[nm], dtype=float))
    # parameters
    params: Dict[str, S.Param] = {}
    triggered: Dict[str, sp.Expr] = {}
    for nm, p in eqs.PARAM.items():
        if getattr(p, "triggerable", False):
            triggered[nm] = _trigger_expression(nm, p, variables)
            continue
        helper = next((nm[:-len(sfx)] for sfx in _MATRIX_HELPER_SUFFIXES if nm.endswith(sfx)), None)
        if helper is not None and helper in eqs.PARAM and int(getattr(eqs.PARAM[helper], "dim", 1)) == 2:
            continue
        mine = _param_from_reference(nm, p, nm in per_scenario)
        params[mine.name] = mine
    # equations
    eqns: List[Eqn] = []
    for nm, e in eqs.EQNs.items():
        tr = _Translator(variables, params, nm, triggered)
        cname = type(e).__name__
        if cname in ("LoopEqn", "LoopOde"):
            used = [k for k in e.var_map if k in triggered]
            if used:
                raise NotImplementedError(f"equation {nm}: triggerable parameters {used} inside a loop body")
            for vm_name, obj in e.var_map.items():           # Set parameters that only the body rewriter registered
                if vm_name not in variables and vm_name not in params:
                    mine = _param_from_reference(vm_name, obj)
                    params[mine.name] = mine
            outer = S.Idx(str(e.outer_index), int(e.n_outer))
            body = tr.loop_body(e.body, e.outer_index, outer, {})
            if cname == "LoopOde":
                dv = e.diff_var
                dvm = tr.loop_body(dv, e.outer_index, outer, {}) if isinstance(dv, sp.Indexed) else tr.expr(dv)
                eqns.append(LoopOde(nm, outer_index=outer, body=body, diff_var=dvm))
            else:
                eqns.append(LoopEqn(nm, outer_index=outer, body=body))
        elif cname == "Ode":
            eqns.append(Ode(nm, tr.expr(e.RHS), diff_var=tr.expr(e.diff_var)))
        elif cname == "Eqn":
            eqns.append(Eqn(nm, tr.expr(e.RHS)))
        else:
            raise NotImplementedError(f"equation {nm}: unsupported equation class {cname}")
    mine = cls(eqns, variables, params, int(getattr(eqs, "nstep", 0) or 0))
    for a in [p for p in params.values() if isinstance(p, S.AliasVar)]:
        if a.value is None or a.size != variables[a.base_var].size:
            a.value = variables[a.base_var].value.copy()
    y = Vars(mine.var_address, np.concatenate([variables[k].value for k in mine.var_address.names])
             if variables else np.zeros(0))
    mine.to_ir()
    return mine, y
