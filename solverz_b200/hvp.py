"""Hessian-vector product of an algebraic model as a second row-template model.

Reference: ``Hvp`` (Solverz/equation/hvp.py:12-74) multiplies every Jacobian block by the matching slice of an
internal parameter ``v_`` -- ``eqn_column[e] = sum_var dF_e/dvar * v_[addr(var)]`` (:19-53) -- and differentiates
that column again (:55-72); ``HVP(y, p, v)`` returns the sparse matrix ``d(J(y) v)/dy``.

Here the column ``G(y; v) = J(y) v`` is written back into the same IR the equations use: every variable load gets
a twin load of the per-scenario parameter ``v__<var>`` at the same index, every walker sum ``S_w`` a twin walker
``sum_kk  sum_z d inner_w/dz * v_z`` over the same CSR row, and the body becomes
``sum_s dF/ds * v_s + sum_w dF/dS_w * S'_w``.  The Jacobian of that model with respect to ``y`` -- derived, frozen
and printed by exactly the machinery of the first derivative -- is the Hessian-vector product, evaluated by the
same kind of one-thread-per-row CUDA kernel."""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import sympy as sp

from .equations import EqTemplate, Load, ModelIR, ParamIR, Walker

__all__ = ["hvp_ir", "V_PREFIX"]

V_PREFIX = "v__"


def hvp_ir(ir: ModelIR) -> ModelIR:
    if ir.kind != "AE":
        raise NotImplementedError("make_hvp=True: the Hessian-vector product is defined for algebraic models (AE)")
    params: Dict[str, ParamIR] = dict(ir.params)
    for name, (_, size) in ir.var_off.items():
        params[V_PREFIX + name] = ParamIR(V_PREFIX + name, "f64", size, True, np.zeros(size))
    eqs: List[EqTemplate] = []
    for eq in ir.eqs:
        loads = dict(eq.loads)
        twin: Dict[sp.Symbol, sp.Symbol] = {}
        nxt = 1 + max((int(s.name[2:]) for s in eq.loads), default=-1)     # loads are numbered `_L<k>`
        for s, ld in eq.loads.items():
            if ld.kind == "var":
                vs = sp.Symbol(f"_L{nxt + len(twin)}", real=True)
                twin[s] = vs
                loads[vs] = Load("par", V_PREFIX + ld.base, ld.index, ld.walker)
        walkers = list(eq.walkers)
        body = sp.Integer(0)
        for s, vs in twin.items():
            if eq.loads[s].index.kind != "inner":
                body += sp.diff(eq.expr, s) * vs
        for w in eq.walkers:
            dS = sp.diff(eq.expr, w.sym)
            if dS == 0:
                continue
            inner = sp.Integer(0)
            for s, vs in twin.items():
                d = sp.diff(w.expr, s)
                if d != 0:
                    inner += d * vs
            if inner == 0:
                continue
            # the twin walker gets its own copies of the per-non-zero loads (the Jacobian derivation and the
            # printer attribute an inner load to exactly one walker); both walk the same CSR row in one loop
            # (jacobian.walker_groups), the compiler merges the duplicate reads
            wid = len(walkers)
            ren = {}
            for s in sorted(inner.free_symbols, key=lambda x: x.name):
                ld = loads.get(s)
                if ld is not None and (ld.index.kind == "inner" or ld.kind == "mat"):
                    s2 = sp.Symbol(f"_L{1 + max(int(k.name[2:]) for k in loads)}", real=True)
                    loads[s2] = Load(ld.kind, ld.base, ld.index, wid)
                    ren[s] = s2
            wsym = sp.Symbol(f"_S{wid}", real=True)
            walkers.append(Walker(w.matrix, w.row, inner.xreplace(ren), wsym))
            body += dS * wsym
        eqs.append(EqTemplate(eq.name, eq.row0, eq.n_rows, body, loads, walkers))
    return ModelIR("AE", ir.n, dict(ir.var_off), params, dict(ir.matrices), eqs, None, 0)


class HvpEquations:
    """Duck-type of ``Equations`` for ``render_cuda_module``: an IR and nothing else."""

    def __init__(self, ir: ModelIR):
        self._ir = hvp_ir(ir)
        self.kind = "AE"

    def to_ir(self) -> ModelIR:
        return self._ir
