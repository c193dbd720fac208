"""Equations, models and their lowering to the row-template IR the CUDA printer consumes.

Reference counterparts (host-only; the IR is what this backend adds):
``Eqn`` / ``Ode`` (Solverz/equation/eqn.py:21-105, 2202-2220), ``LoopEqn`` / ``LoopOde``
(eqn.py:890-1620), ``AE`` / ``DAE`` / ``FDAE`` (equation/equations.py), ``Model.create_instance``
(model/basic.py:36-128).  Ordering rules kept: variables in declaration order; equations in
declaration order, and for a DAE all ``Ode`` before all ``Eqn`` (equations.py:672).

IR: every equation is ``for i in range(n_rows): F[row0 + i] = body(i)`` where ``body`` is a scalar
sympy tree over *loads* (``Y[...]`` / parameter reads whose index is affine in ``i``, gathered
through an int array, or the column index of a CSR row walk) and *walker sums*
(``sum_{kk in row r(i)} inner(kk)``).  Vector equations, ``Mat_Mul`` with a frozen matrix and
``LoopEqn`` bodies all lower to that single form.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import scipy.sparse as sps
import sympy as sp

from . import sym as S
from .sym import AliasVar, Idx, Param, Set, TimeSeriesParam, Var
from .variables import Address, Vars

__all__ = ["Eqn", "Ode", "LoopEqn", "LoopOde", "Model", "Equations", "AE", "DAE", "FDAE", "ModelIR",
           "EqTemplate", "Load", "Index", "Walker", "ParamIR"]


# --------------------------------------------------------------------------------------
# user-facing equation objects
# --------------------------------------------------------------------------------------
class Eqn:
    """``0 = expr`` (vector or scalar)."""

    def __init__(self, name: str, eqn):
        self.name = name
        self.RHS = sp.sympify(eqn)
        self.outer: Optional[Idx] = None


class Ode(Eqn):
    """``d(diff_var)/dt = f``."""

    def __init__(self, name: str, f, diff_var):
        super().__init__(name, f)
        self.diff_var = sp.sympify(diff_var)


class LoopEqn(Eqn):
    """One templated scalar body over an outer index (reference eqn.py:890-1620)."""

    def __init__(self, name: str, outer_index: Idx = None, body=None, model=None, outer: Set = None):
        if outer_index is None and outer is not None:
            raise TypeError("pass outer_index=set.idx('i') (the body must be written with that index)")
        if not isinstance(outer_index, Idx):
            raise TypeError("LoopEqn: outer_index must be an Idx")
        super().__init__(name, body)
        self.outer = outer_index


class LoopOde(LoopEqn):
    def __init__(self, name: str, outer_index: Idx = None, body=None, diff_var=None, model=None):
        super().__init__(name, outer_index, body, model)
        if diff_var is None:
            raise TypeError("LoopOde needs diff_var")
        self.diff_var = sp.sympify(diff_var)


# --------------------------------------------------------------------------------------
# IR
# --------------------------------------------------------------------------------------
@dataclasses.dataclass(frozen=True)
class Index:
    kind: str          # 'affine' : a*i + b | 'gather' : array[a*i + b] | 'inner' : walker column j
    a: int = 1
    b: int = 0
    array: str = ""    # name of the int32 parameter for 'gather'

    def eval(self, i: np.ndarray, arrays: Dict[str, np.ndarray]) -> np.ndarray:
        if self.kind == "affine":
            return self.a * i + self.b
        if self.kind == "gather":
            return arrays[self.array][self.a * i + self.b]
        raise ValueError("inner index has no closed form")


@dataclasses.dataclass(frozen=True)
class Load:
    kind: str          # 'var' | 'par' | 'mat' (data of the enclosing walker's matrix) | 'time'
    base: str
    index: Index
    walker: int = -1   # for index.kind == 'inner' and kind == 'mat'


@dataclasses.dataclass
class Walker:
    matrix: str
    row: Index
    expr: sp.Expr
    sym: sp.Symbol


@dataclasses.dataclass
class EqTemplate:
    name: str
    row0: int
    n_rows: int
    expr: sp.Expr
    loads: Dict[sp.Symbol, Load]
    walkers: List[Walker]
    is_ode: bool = False
    diff_base: str = ""
    diff_index: Optional[Index] = None


@dataclasses.dataclass
class ParamIR:
    name: str
    dtype: str                 # 'f64' | 'i32'
    length: int
    per_scenario: bool
    value: np.ndarray          # default value ([length])
    ts: Optional[Tuple[str, str, Optional[int]]] = None  # (times param, values param, index)
    pattern_of: str = ""       # for matrix index arrays: canonical matrix with the same pattern


@dataclasses.dataclass
class ModelIR:
    kind: str                                  # 'AE' | 'DAE' | 'FDAE'
    n: int
    var_off: Dict[str, Tuple[int, int]]        # name -> (offset, size)
    params: Dict[str, ParamIR]                 # ordered = parameter table order
    matrices: Dict[str, sps.csr_matrix]
    eqs: List[EqTemplate]
    mass: Optional[Tuple[np.ndarray, np.ndarray]] = None  # DAE: rows, cols of the unit entries of M
    nstep: int = 0

    def int_arrays(self) -> Dict[str, np.ndarray]:
        return {k: p.value for k, p in self.params.items() if p.dtype == "i32"}


class _Lowerer:
    def __init__(self, model: "Equations", eq: Eqn, row0: int):
        self.m = model
        self.eq = eq
        self.row0 = row0
        self.loads: Dict[sp.Symbol, Load] = {}
        self.by_key: Dict[Load, sp.Symbol] = {}
        self.walkers: List[Walker] = []
        self.lengths: List[int] = []

    def _load_sym(self, ld: Load) -> sp.Symbol:
        s = self.by_key.get(ld)
        if s is None:
            s = sp.Symbol(f"_L{len(self.by_key)}", real=True)
            self.by_key[ld] = s
            self.loads[s] = ld
        return s

    def _idx_desc(self, key, outer: Optional[Idx], inner: Optional[Idx], size: int, who: str) -> Index:
        if key is None:
            if inner is not None and outer is None and size > 1:
                return Index("inner")          # whole vector inside Mat_Mul
            if size > 1:
                self.lengths.append(size)
                return Index("affine", 1, 0)
            return Index("affine", 0, 0)
        if isinstance(key, int):
            k = key + size if key < 0 else key
            if not 0 <= k < size:
                raise IndexError(f"{who}: index {key} out of range for size {size}")
            return Index("affine", 0, k)
        if isinstance(key, slice):
            if outer is not None:
                raise TypeError(f"{who}: slices are not allowed inside a LoopEqn body")
            start, stop, step = key.indices(size)
            ln = len(range(start, stop, step))
            if ln > 1:
                self.lengths.append(ln)
                return Index("affine", step, start)
            return Index("affine", 0, start)
        if isinstance(key, Idx):
            if key is inner or (inner is not None and key.name == inner.name and key is not outer):
                if key.set is not None and not key.set.is_identity:
                    raise NotImplementedError(f"{who}: summation index over a subset Set is not supported")
                return Index("inner")
            if outer is None or not (key is outer or key.name == outer.name):
                raise ValueError(f"{who}: index {key.name!r} is neither the outer index nor bound by a Sum")
            if key.set is not None and not key.set.is_identity:
                self.m._need_int_array(key.set.name, key.set.values)
                return Index("gather", 1, 0, key.set.name)
            return Index("affine", 1, 0)
        if isinstance(key, tuple) and key[0] == "gather":
            _, par, idx = key
            inner_d = self._idx_desc(idx, outer, inner, par.size, who)
            if inner_d.kind != "affine":
                raise NotImplementedError(f"{who}: nested gathers are not supported")
            self.m._need_int_array(par.name, par.value)
            return Index("gather", inner_d.a, inner_d.b, par.name)
        raise TypeError(f"{who}: unsupported index {key!r}")

    def lower(self, expr: sp.Expr, outer: Optional[Idx], inner: Optional[Idx], wid: int = -1) -> sp.Expr:
        repl = {}
        for s in sorted(expr.free_symbols, key=lambda x: x.name):
            if s is S.t or s.name == "t":
                repl[s] = self._load_sym(Load("time", "t", Index("affine", 0, 0)))
                continue
            si = S.sum_info(s)
            if si is not None:
                if inner is not None:
                    raise NotImplementedError(f"{self.eq.name}: nested Sum / Mat_Mul is not supported")
                repl[s] = self._lower_sum(si, outer)
                continue
            li = S.leaf_info(s)
            if li is None:
                raise ValueError(f"{self.eq.name}: unknown symbol {s} (declare it as Var / Param)")
            base = self.m._resolve(li.base)
            who = f"{self.eq.name}: {s.name}"
            if isinstance(base, Var):
                d = self._idx_desc(li.key, outer, inner, base.size, who)
                repl[s] = self._load_sym(Load("var", base.name, d, wid if d.kind == "inner" else -1))
            elif isinstance(base, Param) and base.dim == 2:
                raise ValueError(f"{who}: a 2-D parameter may only appear as M[i, j] inside Sum(..., j) "
                                 f"or as the matrix of Mat_Mul")
            else:
                d = self._idx_desc(li.key, outer, inner, base.size, who)
                self.m._need_param(base)
                repl[s] = self._load_sym(Load("par", base.name, d, wid if d.kind == "inner" else -1))
        return expr.xreplace(repl)

    def _lower_sum(self, si: S.SumInfo, outer: Optional[Idx]) -> sp.Symbol:
        wid = len(self.walkers)
        wsym = sp.Symbol(f"_S{wid}", real=True)
        if si.matrix is not None:  # Mat_Mul(A, x): row = outer position, column index = inner
            if outer is not None:
                raise NotImplementedError("Mat_Mul inside a LoopEqn body")
            A = self.m._resolve(si.matrix)
            self.m._need_matrix(A)
            self.lengths.append(A.value.shape[0])
            dummy = Idx("_mm_j", A.value.shape[1])
            self.walkers.append(None)  # reserve
            inner_expr = self.lower(si.expr, None, dummy, wid)
            data = self._load_sym(Load("mat", A.name, Index("inner"), wid))
            self.walkers[wid] = Walker(A.name, Index("affine", 1, 0), data * inner_expr, wsym)
            return wsym
        # Sum(expr, j): find the sparse matrix element M[row, j] that defines the walk
        j = si.idx
        mats = []
        for s in si.expr.free_symbols:
            li = S.leaf_info(s)
            if li is not None and isinstance(li.key, tuple) and len(li.key) == 2 and li.key[0] != "gather":
                mats.append((s, li))
        if len(mats) != 1:
            raise NotImplementedError(
                f"{self.eq.name}: each Sum must contain exactly one 2-D parameter element M[i, j] "
                f"(multi-walker Sums are rejected by the reference too, eqn.py:1299); got {len(mats)}")
        msym, mli = mats[0]
        A = self.m._resolve(mli.base)
        rk, ck = mli.key
        if not (isinstance(ck, Idx) and ck.name == j.name):
            raise NotImplementedError(f"{self.eq.name}: the column index of {msym} must be the Sum index")
        self.m._need_matrix(A)
        row = self._idx_desc(rk, outer, None, A.value.shape[0], f"{self.eq.name}: {msym}")
        self.walkers.append(None)
        data = self._load_sym(Load("mat", A.name, Index("inner"), wid))
        if sp.diff(si.expr, msym, 2) != 0:
            raise NotImplementedError(f"{self.eq.name}: Sum body must be linear in {msym}")
        lin = sp.diff(si.expr, msym)
        const = si.expr.xreplace({msym: sp.Integer(0)})
        if const != 0:
            raise NotImplementedError(f"{self.eq.name}: Sum body must be a multiple of {msym} "
                                      f"(terms without the matrix element would be summed over a dense range)")
        inner_expr = data * self.lower(lin, outer, j, wid)
        self.walkers[wid] = Walker(A.name, row, inner_expr, wsym)
        return wsym


class Equations:
    """Container of a lowered model: what ``Model.create_instance`` returns next to ``y0``."""

    kind = "AE"

    def __init__(self, eqns: Sequence[Eqn], variables: Dict[str, Var], params: Dict[str, Param],
                 nstep: int = 0):
        self.EQNs = {e.name: e for e in eqns}
        if len(self.EQNs) != len(eqns):
            raise ValueError("duplicate equation names")
        self.vars = variables
        self.PARAM = params
        self.nstep = nstep
        self.var_address = Address()
        for name, v in variables.items():
            if v.value is None:
                raise ValueError(f"Variable {name} not initialized!")
            self.var_address.add(name, v.size)
        self.a = Address()  # equation addresses, filled by to_ir()
        self._ir: Optional[ModelIR] = None
        self._params_ir: Dict[str, ParamIR] = {}
        self._matrices: Dict[str, sps.csr_matrix] = {}

    # --- helpers used by the lowerer
    def _resolve(self, base):
        if isinstance(base, Var):
            mine = self.vars.get(base.name)
        else:
            mine = self.PARAM.get(base.name)
        if mine is None:
            raise ValueError(f"symbol {base.name!r} is used in an equation but not declared in the model")
        return mine

    def _need_int_array(self, name: str, values: np.ndarray):
        if name not in self._params_ir:
            self._params_ir[name] = ParamIR(name, "i32", len(values), False, np.asarray(values, dtype=np.int32))

    def _need_param(self, p: Param):
        if p.name in self._params_ir:
            return
        if isinstance(p, TimeSeriesParam):
            tn, vn = p.name + "__ts_t", p.name + "__ts_v"
            self._params_ir[p.name] = ParamIR(p.name, "f64", p.size, False, p.value.astype(float),
                                              ts=(tn, vn, p.index))
            self._params_ir[tn] = ParamIR(tn, "f64", len(p.time_series), p.ts_per_scenario, p.time_series)
            self._params_ir[vn] = ParamIR(vn, "f64", len(p.v_series), p.ts_per_scenario, p.v_series)
            return
        self._params_ir[p.name] = ParamIR(p.name, "i32" if p.is_int else "f64", p.size, p.per_scenario,
                                          p.value.astype(np.int32) if p.is_int else p.value.astype(float))

    def _need_matrix(self, A: Param):
        if A.name in self._matrices:
            return
        M = A.value
        self._matrices[A.name] = M
        canon = A.name
        for other, Mo in self._matrices.items():
            if other != A.name and Mo.shape == M.shape and np.array_equal(Mo.indptr, M.indptr) \
                    and np.array_equal(Mo.indices, M.indices):
                canon = self._params_ir[other + "__indptr"].pattern_of or other
                break
        self._params_ir[A.name + "__indptr"] = ParamIR(A.name + "__indptr", "i32", len(M.indptr), False,
                                                       M.indptr.astype(np.int32), pattern_of=canon)
        self._params_ir[A.name + "__indices"] = ParamIR(A.name + "__indices", "i32", len(M.indices), False,
                                                        M.indices.astype(np.int32), pattern_of=canon)
        self._params_ir[A.name + "__data"] = ParamIR(A.name + "__data", "f64", len(M.data), False,
                                                     M.data.astype(float))

    # --- lowering
    def _ordered_eqns(self) -> List[Eqn]:
        return list(self.EQNs.values())

    def to_ir(self) -> ModelIR:
        if self._ir is not None:
            return self._ir
        var_off = {name: (self.var_address[name].start, self.var_address[name].stop - self.var_address[name].start)
                   for name in self.var_address.names}
        n = self.var_address.total_size
        eqs: List[EqTemplate] = []
        row0 = 0
        mass_r: List[np.ndarray] = []
        mass_c: List[np.ndarray] = []
        self.a = Address()
        for e in self._ordered_eqns():
            lw = _Lowerer(self, e, row0)
            if e.outer is not None:
                expr = lw.lower(e.RHS, e.outer, None)
                n_rows = e.outer.n
            else:
                expr = lw.lower(e.RHS, None, None)
                ls = set(lw.lengths)
                if len(ls) > 1:
                    raise ValueError(f"Equation {e.name}: incompatible operand lengths {sorted(ls)}")
                n_rows = ls.pop() if ls else 1
            tpl = EqTemplate(e.name, row0, n_rows, expr, lw.loads, lw.walkers)
            if isinstance(e, (Ode, LoopOde)):
                dl = _Lowerer(self, e, row0)
                dexpr = dl.lower(e.diff_var, e.outer, None)
                if not (isinstance(dexpr, sp.Symbol) and dexpr in dl.loads and dl.loads[dexpr].kind == "var"):
                    raise ValueError(f"Ode {e.name}: diff_var must be a variable (or a slice / element of one)")
                d = dl.loads[dexpr]
                if e.outer is None and dl.lengths and dl.lengths[0] != n_rows:
                    if n_rows == 1:
                        n_rows = dl.lengths[0]
                        tpl.n_rows = n_rows
                    else:
                        raise ValueError(f"Ode {e.name}: diff_var length {dl.lengths[0]} != equation length {n_rows}")
                tpl.is_ode, tpl.diff_base, tpl.diff_index = True, d.base, d.index
                ii = np.arange(n_rows)
                mass_r.append(row0 + ii)
                mass_c.append(var_off[d.base][0] + d.index.eval(ii, self._int_values()))
            eqs.append(tpl)
            self.a.add(e.name, n_rows)
            row0 += n_rows
        if row0 != n:
            raise ValueError(f"Equation size {row0} and variable size {n} not equal!")
        mass = None
        if self.kind == "DAE":
            mass = (np.concatenate(mass_r) if mass_r else np.zeros(0, int),
                    np.concatenate(mass_c) if mass_c else np.zeros(0, int))
        self._ir = ModelIR(self.kind, n, var_off, dict(self._params_ir), dict(self._matrices), eqs, mass,
                           self.nstep)
        return self._ir

    def _int_values(self) -> Dict[str, np.ndarray]:
        return {k: p.value for k, p in self._params_ir.items() if p.dtype == "i32"}

    @property
    def vsize(self):
        return self.var_address.total_size

    @property
    def eqn_size(self):
        return self.to_ir().n

    @property
    def M(self):
        """Singular mass matrix of ones (reference: DAE.M, equations.py:851-902)."""
        ir = self.to_ir()
        if ir.mass is None:
            raise AttributeError("only a DAE has a mass matrix")
        r, c = ir.mass
        return sps.csc_array((np.ones(len(r)), (r, c)), shape=(ir.n, ir.n))


class AE(Equations):
    kind = "AE"


class FDAE(Equations):
    kind = "FDAE"


class DAE(Equations):
    kind = "DAE"

    def _ordered_eqns(self) -> List[Eqn]:
        odes = [e for e in self.EQNs.values() if isinstance(e, (Ode, LoopOde))]
        alg = [e for e in self.EQNs.values() if not isinstance(e, (Ode, LoopOde))]
        return odes + alg


class Model:
    """Attribute bag, like the reference's ``Model`` (model/basic.py:15-128)."""

    def create_instance(self) -> Tuple[Equations, Vars]:
        eqns: List[Eqn] = []
        variables: Dict[str, Var] = {}
        params: Dict[str, Param] = {}
        for key, value in vars(self).items():
            if isinstance(value, Eqn):
                eqns.append(value)
            elif isinstance(value, Var):
                variables[value.name] = value
            elif isinstance(value, Param):
                params[value.name] = value
        is_dae = any(isinstance(e, (Ode, LoopOde)) for e in eqns)
        aliases = [p for p in params.values() if isinstance(p, AliasVar)]
        nstep = max((a.step for a in aliases), default=0)
        if is_dae and aliases:
            raise ValueError("DAE object cannot have iAliasVar!")      # model/basic.py:70-72
        for a in aliases:
            if a.base_var not in variables:
                raise ValueError(f"AliasVar of unknown variable {a.base_var!r}")
            if a.value is None or a.size != variables[a.base_var].size:
                a.value = variables[a.base_var].value.copy()
        cls = DAE if is_dae else (FDAE if aliases else AE)
        eqs = cls(eqns, variables, params, nstep)
        y0 = Vars(eqs.var_address, np.concatenate([variables[k].value for k in eqs.var_address.names])
                  if variables else np.zeros(0))
        eqs.to_ir()
        return eqs, y0
