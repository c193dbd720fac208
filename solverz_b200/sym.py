"""Symbols of the modelling front-end accepted by the CUDA printer.

The reference's modelling DSL (``Solverz/variable/ssymbol.py``, ``equation/param.py``,
``equation/eqn.py:473-560`` for ``Set``/``Idx``/``Sum``, ``sym_algebra/functions.py``) is a host-only
symbolic layer and is out of scope as a subsystem.  What the hot path
needs from it is its *vocabulary*: the leaves and functions that can occur in the trees of
``F(t,y,p)``.  This module provides exactly that vocabulary on top of plain sympy so that models
can be written with the same names (``Var``, ``Param``, ``TimeSeriesParam``, ``Set``, ``Idx``,
``Sum``, ``Mat_Mul``, ``sin`` ...) without importing the reference, and so that
``solverz_b200.adapters.from_solverz`` can lower real Solverz objects into the same trees.

Every leaf is an ordinary ``sympy.Symbol`` whose meaning (base array + index rule) is kept in a
registry; sympy therefore differentiates, simplifies and CSEs the trees without knowing
anything about arrays.
"""
from __future__ import annotations

import itertools
from typing import Any, Dict, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.sparse as sps
import sympy as sp

# --------------------------------------------------------------------------------------
# registry of leaf meanings
# --------------------------------------------------------------------------------------
_LEAVES: Dict[str, "LeafInfo"] = {}
_SUMS: Dict[str, "SumInfo"] = {}
_counter = itertools.count()

t = sp.Symbol("t", real=True)  # time, for DAE / FDAE right-hand sides


class LeafInfo:
    __slots__ = ("base", "key")

    def __init__(self, base, key):
        self.base = base  # Var | Param
        self.key = key    # None (whole) | int | slice | Idx | ('gather', Param, Idx) | (row, col)


class SumInfo:
    __slots__ = ("expr", "idx", "matrix")

    def __init__(self, expr, idx, matrix=None):
        self.expr = expr      # inner expression
        self.idx = idx        # the summation Idx (None for Mat_Mul)
        self.matrix = matrix  # Param (2-D) for Mat_Mul


def leaf_info(sym: sp.Symbol) -> Optional[LeafInfo]:
    return _LEAVES.get(sym.name)


def sum_info(sym: sp.Symbol) -> Optional[SumInfo]:
    return _SUMS.get(sym.name)


def _leaf(base, key, keyrepr: str) -> sp.Symbol:
    name = base.name if key is None else f"{base.name}[{keyrepr}]"
    old = _LEAVES.get(name)
    if old is not None and old.base is not base:
        # a new model re-using a name: the newest declaration wins for new expressions
        pass
    _LEAVES[name] = LeafInfo(base, key)
    return sp.Symbol(name, real=True)


class _ExprMixin:
    """Arithmetic on front-end objects is delegated to their sympy symbol."""

    def _sympy_(self):
        return self.sym

    def __add__(self, o): return self.sym + o
    def __radd__(self, o): return o + self.sym
    def __sub__(self, o): return self.sym - o
    def __rsub__(self, o): return o - self.sym
    def __mul__(self, o): return self.sym * o
    def __rmul__(self, o): return o * self.sym
    def __truediv__(self, o): return self.sym / o
    def __rtruediv__(self, o): return o / self.sym
    def __pow__(self, o): return self.sym ** o
    def __rpow__(self, o): return o ** self.sym
    def __neg__(self): return -self.sym
    def __pos__(self): return self.sym


class Idx:
    """Bounded loop index (reference: ``sympy.Idx`` as used by LoopEqn, eqn.py:890-960)."""

    def __init__(self, name: str, n: Optional[int] = None, set_: Optional["Set"] = None):
        self.name = name
        self.set = set_
        self.n = int(n) if n is not None else (len(set_) if set_ is not None else None)
        if self.n is None:
            raise ValueError("Idx needs a length or a Set")

    def __repr__(self):
        return self.name


class Set:
    """Named index set (reference: ``IndexSet``, eqn.py:473-560).

    ``Set('Bus', nb)`` is the identity range; ``Set('PVPQ', values)`` an explicit subset whose
    values must be non-negative and duplicate-free.
    """

    def __init__(self, name: str, values: Union[int, Sequence[int]]):
        self.name = name
        if isinstance(values, (int, np.integer)):
            self.values = np.arange(int(values), dtype=np.int64)
            self.is_identity = True
        else:
            v = np.asarray(list(values) if isinstance(values, range) else values)
            if v.ndim != 1 or (v.size and not np.issubdtype(v.dtype, np.integer)):
                raise TypeError(f"Set {name!r}: values must be a 1-D integer sequence")
            if v.size and v.min() < 0:
                raise ValueError(f"Set {name!r}: values must be non-negative")
            if len(np.unique(v)) != len(v):
                raise ValueError(f"Set {name!r}: values must be duplicate-free")
            self.values = v.astype(np.int64)
            self.is_identity = bool(v.size and np.array_equal(v, np.arange(v.size)))

    def __len__(self):
        return len(self.values)

    def idx(self, name: str) -> Idx:
        return Idx(name, set_=self)


def _key_repr(key) -> str:
    if isinstance(key, slice):
        return f"{key.start if key.start is not None else ''}:{key.stop if key.stop is not None else ''}" + (
            f":{key.step}" if key.step not in (None, 1) else "")
    if isinstance(key, Idx):
        return key.name if (key.set is None or key.set.is_identity) else f"{key.set.name}[{key.name}]"
    if isinstance(key, tuple) and len(key) == 3 and key[0] == "gather":
        return f"{key[1].name}[{_key_repr(key[2])}]"
    if isinstance(key, tuple):
        return ", ".join(_key_repr(k) for k in key)
    return str(key)


class _Indexable(_ExprMixin):
    name: str

    def _index_key(self, key):
        if isinstance(key, (int, np.integer)):
            return int(key)
        if isinstance(key, (slice, Idx)):
            return key
        if isinstance(key, sp.Symbol):  # m.Vm[m.pv_pq_idx[i]] : explicit indirect index
            li = leaf_info(key)
            if li is not None and isinstance(li.base, Param) and li.base.is_int and isinstance(li.key, Idx):
                return ("gather", li.base, li.key)
        raise TypeError(f"unsupported index {key!r} on {self.name}; use int, slice, Idx or an int Param "
                        f"indexed by an Idx (list / fancy indices are rejected like in the reference, "
                        f"code_printer/python/utilities.py:257-261)")

    def __getitem__(self, key):
        if isinstance(key, tuple):
            if getattr(self, "dim", 1) != 2 or len(key) != 2:
                raise TypeError(f"{self.name}: 2 indices only on 2-D parameters")
            k = (self._index_key(key[0]), self._index_key(key[1]))
            return _leaf(self, k, _key_repr(k))
        k = self._index_key(key)
        return _leaf(self, k, _key_repr(k))


class Var(_Indexable):
    """Unknown vector (reference: ``Var`` in variable/ssymbol.py)."""

    def __init__(self, name: str, value=None):
        self.name = name
        self.value = None if value is None else np.atleast_1d(np.asarray(value, dtype=float)).ravel()
        self.sym = _leaf(self, None, "")

    @property
    def size(self):
        return 0 if self.value is None else self.value.size

    def __repr__(self):
        return f"Var({self.name})"


class Param(_Indexable):
    """Parameter (reference: ``Param``, equation/param.py:12-115).

    1-D float / int arrays, or a frozen 2-D matrix (``dim=2``; sparse or dense) that is lowered
    to CSR row walkers.  ``per_scenario=True`` marks a 1-D parameter whose value differs between
    the scenarios of a batch; its value then has shape ``[len]`` (a default replicated to every
    scenario) and the batched value ``[B, len]`` is supplied at solve time.
    """

    def __init__(self, name: str, value=None, triggerable: bool = False, trigger_var=None,
                 trigger_fun=None, dim: int = 1, dtype=float, sparse: bool = False,
                 per_scenario: bool = False):
        if triggerable:
            raise NotImplementedError("triggerable parameters are host callbacks and are not supported "
                                      "by the CUDA backend")
        self.name = name
        self.dim = dim
        self.sparse = sparse
        self.is_int = dtype is int or (isinstance(dtype, type) and issubclass(dtype, (int, np.integer)))
        self.per_scenario = per_scenario
        if dim == 2:
            if per_scenario:
                raise NotImplementedError("2-D parameters are frozen at build time (param.py:12-55)")
            self.value = sps.csr_matrix(value)
            self.value.sort_indices()
        elif value is None:
            self.value = None
        else:
            self.value = np.atleast_1d(np.asarray(value, dtype=np.int64 if self.is_int else float)).ravel()
        self.sym = _leaf(self, None, "")

    @property
    def size(self):
        if self.value is None:
            return 0
        return self.value.shape[0] if self.dim == 2 else self.value.size

    def __repr__(self):
        return f"Param({self.name})"


class TimeSeriesParam(Param):
    """Piecewise-linear parameter of time (reference: param.py:120-212).

    Value at ``t``: linear interpolation for ``t < time_series[-1]``, else the LAST value
    (param.py:187-212).  With ``index`` given, the interpolated scalar overwrites entry ``index``
    of the base vector ``value``.  ``per_scenario=True``: the knot arrays are ``[B, nk]``.
    """

    def __init__(self, name, v_series, time_series, value=None, index=None, per_scenario=False):
        v_series = np.asarray(v_series, dtype=float).ravel()
        time_series = np.asarray(time_series, dtype=float).ravel()
        if len(v_series) != len(time_series):
            raise ValueError("Incompatible length between value series and time series!")
        if not np.all(np.diff(time_series) > 0):
            raise ValueError("Time stamp should be strictly monotonically increasing!")
        if value is None:
            value = v_series[0:1]
        super().__init__(name, value, per_scenario=False)
        self.v_series = v_series
        self.time_series = time_series
        self.index = index
        self.ts_per_scenario = per_scenario
        if index is not None and not isinstance(index, (int, np.integer)):
            raise NotImplementedError("TimeSeriesParam: only a scalar index is supported")

    def get_v_t(self, tt):
        if tt is None:
            return self.value
        vt = np.interp(tt, self.time_series, self.v_series) if tt < self.time_series[-1] else self.v_series[-1]
        if self.index is None:
            return np.atleast_1d(vt)
        out = self.value.copy()
        out[self.index] = vt
        return out


class AliasVar(Param):
    """Value of variable ``name`` at an earlier time node (reference: ``AliasVar``,
    variable/ssymbol.py:111-115): ``step=1`` is y_0 (the node the step starts from), ``step=2`` is
    y_{-1}, ...  Lowered to a per-scenario parameter named ``<name>_tag_<step-1>`` that
    ``fdae_solver`` refreshes on the device between steps."""

    def __init__(self, name: str, step: int = 1, value=None, init=None):
        if step < 1:
            raise ValueError("AliasVar: step must be >= 1")
        self.base_var = name
        self.step = int(step)
        super().__init__(f"{name}_tag_{abs(step - 1)}", value, per_scenario=True)


def Sum(expr, idx: Idx) -> sp.Symbol:
    """``Sum(body, j)`` over a bounded index (reference: eqn.py ``Sum``; lowered to CSR row walks,
    eqn.py:1214-1360)."""
    if not isinstance(idx, Idx):
        raise TypeError("Sum(expr, idx): idx must be an Idx")
    name = f"_szsum{next(_counter)}"
    _SUMS[name] = SumInfo(sp.sympify(expr), idx)
    return sp.Symbol(name, real=True)


def Mat_Mul(A: Param, x) -> sp.Symbol:
    """``A @ x`` with a frozen 2-D parameter ``A`` (reference: functions.py:40-163; F side is
    ``SolCF.csc_matvec``, custom_function.py:202-210)."""
    if not isinstance(A, Param) or A.dim != 2:
        raise TypeError("Mat_Mul(A, x): A must be a 2-D Param; general matrix products are not on the "
                        "hot path")
    name = f"_szmm{next(_counter)}"
    _SUMS[name] = SumInfo(sp.sympify(x), None, A)
    return sp.Symbol(name, real=True)


# --------------------------------------------------------------------------------------
# function vocabulary
# --------------------------------------------------------------------------------------
sin, cos, exp, Abs, atan2 = sp.sin, sp.cos, sp.exp, sp.Abs, sp.atan2
ln = sp.log
tan, sqrt = sp.tan, sp.sqrt


class Sign(sp.Function):
    """Reference ``Sign`` (functions.py:336-361): np.sign, treated as a constant when differentiated
    (sympy's own ``sign`` would differentiate to a Dirac delta)."""
    nargs = 1
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)

    def _cudacode(self, printer):
        return f"szb_sign({printer._print(self.args[0])})"



class heaviside(sp.Function):
    """Reference: ``heaviside`` (functions.py:364-389): 1 for x >= 0 else 0; derivative 0."""
    nargs = 1
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)


class Saturation(sp.Function):
    """Reference: ``Saturation(v, vmin, vmax)`` (functions.py:429-470): clip; dS/dv = In(v,vmin,vmax)."""
    nargs = 3
    is_real = True

    def fdiff(self, argindex=1):
        if argindex == 1:
            return In(*self.args)
        return sp.Integer(0)


class In(sp.Function):
    """1 if vmin <= v <= vmax else 0 (functions.py:473-500); derivative 0."""
    nargs = 3
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)


class GreaterThan(sp.Function):
    nargs = 2
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)


class LessThan(sp.Function):
    nargs = 2
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)


class And(sp.Function):
    nargs = 2
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)


class Or(sp.Function):
    nargs = 2
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)


class Not(sp.Function):
    nargs = 1
    is_real = True

    def fdiff(self, argindex=1):
        return sp.Integer(0)


class Min(sp.Function):
    """Reference ``Min(x, y)`` (functions.py:503-530): d/dx = (x <= y), d/dy = (y < x)."""
    nargs = 2
    is_real = True

    def fdiff(self, argindex=1):
        x, y = self.args
        return LessThan(x, y) + (1 - GreaterThan(x, y) - LessThan(x, y)) if argindex == 1 else GreaterThan(x, y)


class switch(sp.Function):
    """Reference ``switch(v_0, ..., v_{k-1}, flag)`` (functions.py:429-434, custom_function.py:71-89):
    ``v_i`` where ``flag == i``, else 0; the derivative switches between the branch derivatives.
    Printed through the same ``_cudacode`` hook user functions use."""
    is_real = True

    def _eval_derivative(self, s):
        return switch(*[a.diff(s) for a in self.args[:-1]], self.args[-1])

    def _cudacode(self, printer):
        flag = printer._print(self.args[-1])
        out = "0.0"
        for i in reversed(range(len(self.args) - 1)):
            out = f"(({flag}) == {i}.0 ? ({printer._print(self.args[i])}) : {out})"
        return out


def AntiWindUp(u, umin, umax, e):
    """Reference ``AntiWindUp`` (functions.py:461-499): the integrator input ``e`` gated off while the
    PI output sits on a limit and the error pushes further into it."""
    zero = sp.Integer(0)
    return e * Not(Or(And(GreaterThan(u, umax), GreaterThan(e, zero)), And(LessThan(u, umin), LessThan(e, zero))))


__all__ = ["Var", "Param", "TimeSeriesParam", "AliasVar", "Set", "Idx", "Sum", "Mat_Mul", "t", "sin", "cos", "exp",
           "ln", "Abs", "atan2", "tan", "sqrt", "Sign", "heaviside", "Saturation", "In", "GreaterThan",
           "LessThan", "And", "Or", "Not", "Min", "switch", "AntiWindUp", "leaf_info", "sum_info"]
