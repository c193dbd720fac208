"""Symbolic Jacobian of the row-template IR and its frozen sparsity pattern.

Reference counterparts: ``Equations.FormJac`` (Solverz/equation/equations.py:167-320),
``JacBlock.ParseSp`` (jac.py:264-328), ``Jac.parse_row_col_data`` (jac.py:106-126) and, for loop
bodies, ``LoopEqnDiff`` / ``compute_loop_jac_sparsity`` / ``build_loop_jac_kernel_source``
(eqn.py:136-393, loop_jac.py:1438, 1865-2040).  The reference evaluates, per structural non-zero,
the *whole* row sums again (O(nnz * row degree)); here one thread
owns one equation row, walks each CSR row once and emits every entry of that Jacobian row.

The pattern is frozen numerically on the host: every emission site gets, per row (and per walker
non-zero), the CSC slot it writes.  Duplicate (row, col) pairs -- e.g. the diagonal term of a
walker meeting the outer-load entry -- are summed in program order by the owning thread: the first
emission of a slot stores, later ones accumulate (slot index stored bit-complemented).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Tuple

import numpy as np
import sympy as sp

from .equations import EqTemplate, Index, Load, ModelIR, Walker

__all__ = ["JacTemplate", "JacPattern", "derive_jacobian"]


@dataclasses.dataclass
class InnerEntry:
    walker: int
    load: Load                 # var load with index.kind == 'inner'
    coef: sp.Expr              # d body / d S_w   (outer-level expression)
    term: sp.Expr              # d inner_w / d z  (per non-zero expression)


@dataclasses.dataclass
class OuterEntry:
    load: Load
    expr: sp.Expr              # over loads, S symbols and accumulator symbols


@dataclasses.dataclass
class JacTemplate:
    eq: EqTemplate
    accs: List[Tuple[int, sp.Symbol, sp.Expr]]   # (walker, symbol, per-nnz expression) extra sums
    outer: List[OuterEntry]
    inner: List[InnerEntry]
    two_pass: bool
    groups: List[List[int]] = dataclasses.field(default_factory=list)  # walkers sharing one CSR walk
    # filled by the pattern pass
    outer_slots: List[np.ndarray] = dataclasses.field(default_factory=list)   # int32[n_rows] each
    inner_slots: List[np.ndarray] = dataclasses.field(default_factory=list)   # int32[nnz walked] each
    inner_off: Dict[int, np.ndarray] = dataclasses.field(default_factory=dict)  # walker -> int32[n_rows]


@dataclasses.dataclass
class JacPattern:
    n: int
    nnz: int
    indptr: np.ndarray        # CSC
    indices: np.ndarray
    templates: List[JacTemplate]
    mass_slots: Optional[np.ndarray] = None   # slots of the unit entries of M (DAE)


def _is_zero(e) -> bool:
    return e == 0


def walker_groups(ir: ModelIR, eq: EqTemplate) -> List[List[int]]:
    """Walkers over matrices with an identical pattern and the same row rule share one loop."""
    keys: Dict[Tuple[str, Index], int] = {}
    groups: List[List[int]] = []
    for wi, w in enumerate(eq.walkers):
        canon = ir.params[w.matrix + "__indptr"].pattern_of or w.matrix
        k = (canon, w.row)
        if k not in keys:
            keys[k] = len(groups)
            groups.append([])
        groups[keys[k]].append(wi)
    return groups


def derive_templates(ir: ModelIR) -> List[JacTemplate]:
    out = []
    for eq in ir.eqs:
        wsyms = [w.sym for w in eq.walkers]
        outer_loads = [(s, ld) for s, ld in eq.loads.items() if ld.kind == "var" and ld.index.kind != "inner"]
        accs: List[Tuple[int, sp.Symbol, sp.Expr]] = []
        outer: List[OuterEntry] = []
        inner: List[InnerEntry] = []
        dS = [sp.diff(eq.expr, w) for w in wsyms]
        for s, ld in outer_loads:
            d = sp.diff(eq.expr, s)
            for wi, w in enumerate(eq.walkers):
                if _is_zero(dS[wi]):
                    continue
                di = sp.diff(w.expr, s)
                if _is_zero(di):
                    continue
                a = sp.Symbol(f"_A{len(accs)}", real=True)
                accs.append((wi, a, di))
                d = d + dS[wi] * a
            if not _is_zero(d):
                outer.append(OuterEntry(ld, d))
        for wi, w in enumerate(eq.walkers):
            if _is_zero(dS[wi]):
                continue
            for s, ld in eq.loads.items():
                if ld.kind == "var" and ld.index.kind == "inner" and ld.walker == wi:
                    di = sp.diff(w.expr, s)
                    if not _is_zero(di):
                        inner.append(InnerEntry(wi, ld, dS[wi], di))
        sum_syms = set(wsyms) | {a for _, a, _ in accs}
        two_pass = any(e.coef.free_symbols & sum_syms for e in inner)
        out.append(JacTemplate(eq, accs, outer, inner, two_pass, walker_groups(ir, eq)))
    return out


def derive_jacobian(ir: ModelIR) -> JacPattern:
    """Templates + frozen CSC pattern + per-emission slot tables."""
    tpls = derive_templates(ir)
    ints = ir.int_arrays()
    n = ir.n
    rows_l, cols_l, seq_l, site_l = [], [], [], []   # flat emissions
    site_meta = []  # (template idx, 'o'|'i', entry idx, length)
    for ti, jt in enumerate(tpls):
        eq = jt.eq
        ii = np.arange(eq.n_rows, dtype=np.int64)
        # walkers: rows walked and offsets
        walk = {}
        for wi, w in enumerate(eq.walkers):
            M = ir.matrices[w.matrix]
            r = w.row.eval(ii, ints)
            cnt = (M.indptr[r + 1] - M.indptr[r]).astype(np.int64)
            off = np.zeros(eq.n_rows + 1, dtype=np.int64)
            off[1:] = np.cumsum(cnt)
            # flat list of (i, q, column)
            rep_i = np.repeat(ii, cnt)
            q = np.arange(off[-1]) - np.repeat(off[:-1], cnt)
            kk = np.repeat(M.indptr[r], cnt) + q
            walk[wi] = (off, rep_i, q, M.indices[kk].astype(np.int64))
            jt.inner_off[wi] = off[:-1].astype(np.int32)
        n_w = len(jt.groups)
        group_of = {wi: gi for gi, g in enumerate(jt.groups) for wi in g}
        for ei, e in enumerate(jt.inner):
            off, rep_i, q, cols = walk[e.walker]
            rows_l.append(eq.row0 + rep_i)
            cols_l.append(ir.var_off[e.load.base][0] + cols)
            # program order inside a row: walker, position in the CSR row, entry
            seq_l.append((group_of[e.walker] * (1 << 40)) + q * (1 << 12) + ei)
            site_meta.append((ti, "i", ei, len(rep_i)))
        for ei, e in enumerate(jt.outer):
            c = ir.var_off[e.load.base][0] + e.load.index.eval(ii, ints)
            if np.any(c < ir.var_off[e.load.base][0]) or np.any(c >= sum(ir.var_off[e.load.base])):
                raise IndexError(f"{eq.name}: index of {e.load.base} out of range")
            rows_l.append(eq.row0 + ii)
            cols_l.append(c.astype(np.int64))
            seq_l.append(np.full(eq.n_rows, n_w * (1 << 40) + ei, dtype=np.int64))
            site_meta.append((ti, "o", ei, eq.n_rows))
    mass_rc = None
    if ir.mass is not None:
        mass_rc = ir.mass
    if rows_l:
        rows = np.concatenate(rows_l)
        cols = np.concatenate(cols_l)
        seq = np.concatenate(seq_l)
    else:
        rows = cols = seq = np.zeros(0, dtype=np.int64)
    allr, allc = rows, cols
    if mass_rc is not None:
        allr = np.concatenate([rows, mass_rc[0]])
        allc = np.concatenate([cols, mass_rc[1]])
    key = allc * n + allr
    uniq, inv = np.unique(key, return_inverse=True)
    nnz = len(uniq)
    ucol = uniq // n
    urow = uniq % n
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, ucol + 1, 1)
    indptr = np.cumsum(indptr)
    slot = inv[:len(rows)]
    # first-writer flags: among the emissions of one slot (all from the same row => same thread),
    # the one earliest in program order stores, the others accumulate
    order = np.lexsort((seq, slot))
    first = np.ones(len(slot), dtype=bool)
    so = slot[order]
    first[order[1:]] = so[1:] != so[:-1]
    enc = np.where(first, slot, ~slot).astype(np.int32)
    pos = 0
    for ti, kind, ei, ln in site_meta:
        seg = enc[pos:pos + ln]
        pos += ln
        if kind == "i":
            tpls[ti].inner_slots.append(seg)
        else:
            tpls[ti].outer_slots.append(seg)
    mass_slots = None
    if mass_rc is not None:
        mass_slots = inv[len(rows):].astype(np.int32)
    return JacPattern(n, nnz, indptr, urow.astype(np.int64), tpls, mass_slots)
