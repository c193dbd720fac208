"""CUDA printer: row-template IR  ->  one .cu translation unit per model.

This is the ``lang='cuda'`` sibling of the reference's python/numba printers
(Solverz/code_printer/python/module/module_printer.py: ``print_F`` 683-781, ``print_inner_F``
784-870, ``print_sub_inner_F`` 911-1047, ``print_J`` 140-250, ``print_inner_J`` 253-458 and the
LoopEqn kernel builders eqn.py:1462-1520, loop_jac.py:1865-2040).  Where those emit one ``@njit``
function per equation plus a Python wrapper that slices ``y_`` and rebuilds a COO->CSC matrix
on every call, this printer emits

* one ``__host__ __device__`` row function per equation for F and one for the Jacobian row,
  with common sub-expressions hoisted by ``sympy.cse`` and ``sin``/``cos`` of the same argument
  fused into ``sincos``;
* ``szbm_F_kernel`` / ``szbm_J_kernel``: one CTA per scenario (grid-stride), threads over equation
  rows, fp64; ``Y[b]`` and the per-scenario parameters are read through the read-only path, the
  scenario-invariant index arrays (CSR of frozen matrices, index sets, slot tables) are shared by
  all CTAs and stay cache resident; F also reduces ``max|F|`` per scenario (NaN-propagating);
  J writes straight into the CSC slots of the frozen pattern -- no COO, no sort, no atomics;
* the ``szbm_*`` C entry points of include/szb_model.h.

The same translation unit compiles as plain C++ with ``-DSZBM_HOST_CHECK`` into a host library
exposing ``szbm_host_F`` / ``szbm_host_J``; that build exists only so the printer can be tested on
machines without a GPU (tests/), it is never loaded by the product.
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Sequence, Set, Tuple

import numpy as np
import sympy as sp
from sympy.printing.c import C99CodePrinter

from ...equations import EqTemplate, Index, Load, ModelIR
from ...jacobian import JacPattern, JacTemplate, derive_jacobian
from ... import sym as S

__all__ = ["print_model", "source_hash"]


class _CPrinter(C99CodePrinter):
    """C printer of one row function.  Custom functions plug in the way ``_numpycode`` does in the
    reference (docs/src/advanced.md "Writing Custom Functions"; num_api/module_parser.py:13-39 injects
    their numerical modules): a sympy ``Function`` subclass supplies ``_cudacode(self, printer) -> str``
    (the C expression of a call) and, optionally, a class attribute ``_cuda_device_code`` holding the
    ``SZB_HD static inline double name(...) {...}`` definition that expression calls; the definition is
    emitted once into the generated translation unit."""

    snippets: Dict[str, str] = {}     # filled while a model is printed (print_model resets it)

    def __init__(self, names: Dict[sp.Symbol, str]):
        super().__init__(dict(user_functions={
            "heaviside": "szb_heaviside", "Saturation": "szb_saturation", "In": "szb_in",
            "GreaterThan": "szb_gt", "LessThan": "szb_lt", "And": "szb_and", "Or": "szb_or",
            "Not": "szb_not", "Min": "szb_min", "Heaviside": "szb_heaviside",
        }))
        self.names = names

    def _print_Symbol(self, expr):
        return self.names.get(expr, expr.name)

    def _print_Function(self, expr):
        if hasattr(expr, "_cudacode"):
            code = getattr(type(expr), "_cuda_device_code", None)
            if code:
                _CPrinter.snippets.setdefault(code.strip(), code.strip())   # keyed by text: classes may share a block
            return expr._cudacode(self)
        return super()._print_Function(expr)

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer and 2 <= abs(int(e)) <= 4:
            bs = self._print(b)
            if not (b.is_Symbol or b.is_Number):
                bs = f"({bs})"
            prod = "*".join([bs] * abs(int(e)))
            return prod if e > 0 else f"1.0/({prod})"
        return super()._print_Pow(expr)

    def _print_Min(self, expr):
        # sympy's C printer dispatches on the class NAME and would treat our two-argument Min (functions.py:502-520)
        # as its own variadic Min (nested fmin of re-built one-argument nodes)
        a, b = expr.args
        return f"szb_min({self._print(a)}, {self._print(b)})"

    def _print_sign(self, expr):
        return f"szb_sign({self._print(expr.args[0])})"

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Float(self, expr):
        return repr(float(expr))


def _ident(name: str) -> str:
    return "".join(ch if (ch.isalnum() or ch == "_") else "_" for ch in name)


class _Emitter:
    def __init__(self):
        self.lines: List[str] = []
        self.ind = 1

    def w(self, s: str = ""):
        self.lines.append(("  " * self.ind + s) if s else "")

    def text(self) -> str:
        return "\n".join(self.lines)


class _RowGen:
    """Shared code generation for one equation template."""

    def __init__(self, ir: ModelIR, eq: EqTemplate, tabs: "_Tables", hoist_trig: bool = True):
        self.ir = ir
        self.eq = eq
        self.tabs = tabs
        self.hoist_trig = hoist_trig
        self.names: Dict[sp.Symbol, str] = {}
        for s, ld in eq.loads.items():
            self.names[s] = s.name.lstrip("_")
        for w in eq.walkers:
            self.names[w.sym] = w.sym.name.lstrip("_")
        self.cse_count = 0
        self.alias: Dict[sp.Symbol, sp.Symbol] = {}
        self.trig_outer: Dict[sp.Symbol, Tuple[sp.Symbol, sp.Symbol]] = {}
        self.trig_inner: Dict[sp.Symbol, Tuple[sp.Symbol, sp.Symbol]] = {}

    def set_groups(self, groups: List[List[int]]):
        """Inner loads of walkers that share one CSR walk read the same address: merge them so
        that sympy's CSE sees e.g. sin(Va_i - Va_j) of the G walk and of the B walk as one."""
        group_of = {wi: gi for gi, g in enumerate(groups) for wi in g}
        rep: Dict[Tuple, sp.Symbol] = {}
        for s, ld in sorted(self.eq.loads.items(), key=lambda kv: int(kv[0].name[2:])):
            if ld.index.kind == "inner" and ld.kind in ("var", "par"):
                k = (ld.kind, ld.base, group_of.get(ld.walker, -1))
                if k in rep:
                    self.alias[s] = rep[k]
                else:
                    rep[k] = s

    def canon(self, e: sp.Expr) -> sp.Expr:
        return e.xreplace(self.alias) if self.alias else e

    # ---- trig hoisting: sin / cos of (outer-indexed V) - (walk-indexed V) by the angle-difference
    # identities from one sincos per ROW (outer value) and a per-scenario table of sin V, cos V that the
    # kernel fills once (one sincos per variable entry instead of one per walked position and equation).
    def hoist(self, e: sp.Expr) -> sp.Expr:
        e = self.canon(e)
        if not (self.tabs.trig_ok and self.hoist_trig):
            return e

        def split(arg):
            """arg = a*Lo + b*Li with a = -b = +-1, Lo outer / Li inner loads of one variable -> (a, Lo, Li)."""
            if not (arg.is_Add and len(arg.args) == 2):
                return None
            terms = []
            for t in arg.args:
                c, sym = t.as_coeff_Mul()
                if sym not in self.eq.loads or c not in (1, -1):
                    return None
                terms.append((int(c), sym))
            (c1, s1), (c2, s2) = terms
            if c1 != -c2:
                return None
            l1, l2 = self.eq.loads[s1], self.eq.loads[s2]
            if l1.kind != "var" or l2.kind != "var" or l1.base != l2.base:
                return None
            in1, in2 = l1.index.kind == "inner", l2.index.kind == "inner"
            if in1 == in2:
                return None
            return (c2, s2, s1) if in1 else (c1, s1, s2)

        def fix(node):
            sp_ = split(node.args[0])
            if sp_ is None:
                return node
            a, lo, li = sp_
            base = self.eq.loads[lo].base
            self.tabs.trig_table(self.ir, base)
            so, co = self.trig_outer.setdefault(lo, (sp.Symbol(f"so{lo.name[2:]}", real=True),
                                                     sp.Symbol(f"co{lo.name[2:]}", real=True)))
            si, ci = self.trig_inner.setdefault(li, (sp.Symbol(f"si{li.name[2:]}", real=True),
                                                     sp.Symbol(f"ci{li.name[2:]}", real=True)))
            for x in (so, co, si, ci):
                self.names[x] = x.name
            if isinstance(node, sp.sin):                  # sin(a (Lo - Li)) = a (sin Lo cos Li - cos Lo sin Li)
                return a * (so * ci - co * si)
            return co * ci + so * si                      # cos is even
        return e.replace(lambda x: isinstance(x, (sp.sin, sp.cos)), fix)

    def emit_outer_trig(self, em: _Emitter):
        for lo, (so, co) in sorted(self.trig_outer.items(), key=lambda kv: int(kv[0].name[2:])):
            em.w(f"double {so.name}, {co.name};")
            em.w(f"sincos({self.names[lo]}, &{so.name}, &{co.name});")

    def trig_loads(self, exprs: Sequence[sp.Expr]) -> List[Tuple[str, str, str]]:
        """(ctype, name, text) of the table reads the walk body needs."""
        free = set().union(*[e.free_symbols for e in exprs]) if exprs else set()
        out = []
        for li, (si, ci) in sorted(self.trig_inner.items(), key=lambda kv: int(kv[0].name[2:])):
            ld = self.eq.loads[li]
            ix = self.idx_text(ld.index)
            if si in free:
                out.append(("double", si.name, f"c.ts_{_ident(ld.base)}[{ix}]"))
            if ci in free:
                out.append(("double", ci.name, f"c.tc_{_ident(ld.base)}[{ix}]"))
        return out

    # ---- index / load text
    def idx_text(self, ix: Index) -> str:
        if ix.kind == "inner":
            return "j"
        lin = "i" if (ix.a == 1 and ix.b == 0) else (str(ix.b) if ix.a == 0 else f"({ix.a}*i + {ix.b})")
        if ix.kind == "affine":
            return lin
        return f"c.p_{_ident(ix.array)}[{lin}]"

    def load_text(self, ld: Load) -> str:
        if ld.kind == "time":
            return "c.t"
        if ld.kind == "var":
            off = self.ir.var_off[ld.base][0]
            return f"c.Y[{off} + {self.idx_text(ld.index)}]"
        if ld.kind == "mat":
            return f"c.p_{_ident(ld.base)}__data[kk]"
        p = self.ir.params[ld.base]
        ix = self.idx_text(ld.index)
        if p.ts is not None:
            tn, vn, tsi = p.ts
            nk = self.ir.params[tn].length
            ts = f"szb_interp(c.p_{_ident(tn)}, c.p_{_ident(vn)}, {nk}, c.t)"
            if tsi is None:
                return ts
            return f"(({ix}) == {int(tsi)} ? {ts} : c.p_{_ident(ld.base)}[{ix}])"
        if p.dtype == "i32":
            return f"(double)c.p_{_ident(ld.base)}[{ix}]"
        return f"c.p_{_ident(ld.base)}[{ix}]"

    def emit_loads(self, em: _Emitter, syms: Sequence[sp.Symbol], inner: bool):
        for s in sorted(syms, key=lambda x: int(x.name[2:])):
            ld = self.eq.loads[s]
            is_inner = ld.index.kind == "inner" or ld.kind == "mat"
            if is_inner != inner:
                continue
            em.w(f"const double {self.names[s]} = {self.load_text(ld)};")

    def open_walk(self, em: _Emitter, gi: int, canon: str, syms: Sequence[sp.Symbol],
                  extra: Sequence[Tuple[str, str, str]] = (), pipelined: bool = True):
        """Opens the CSR row walk ``for kk in [kb, ke)`` software-pipelined by one iteration: the column
        index, the inner loads (matrix values, gathered variables / parameters) and ``extra`` per-position
        loads ``(ctype, name, text in kk)`` of iteration kk + 1 are issued before iteration kk computes, so
        that the two dependent round trips of a gather (index, then value) overlap the arithmetic (fp64
        ``sincos`` is ~100 instructions) instead of sitting in front of it."""
        import re
        loads: List[Tuple[str, str, str]] = []
        for s_ in sorted(syms, key=lambda x: int(x.name[2:])):
            ld = self.eq.loads[s_]
            if ld.index.kind == "inner" or ld.kind == "mat":
                loads.append(("double", self.names[s_], self.load_text(ld)))
        loads += list(extra)
        if not pipelined:        # plain walk (the F kernel is issue-bound: the extra moves cost more than they hide)
            em.w("{")
            em.w(f"for (int kk = kb{gi}; kk < ke{gi}; ++kk) {{")
            em.ind += 1
            em.w(f"const int j = c.p_{canon}__indices[kk];")
            for cty, nm, tx in loads:
                em.w(f"const {cty} {nm} = {tx};")
            em.w("(void)j;")
            return
        jn = f"jn{gi}"

        def at(text: str, kk_expr: str) -> str:
            return re.sub(r"\bj\b", jn, re.sub(r"\bkk\b", kk_expr, text))

        em.w("{")                                      # own scope: a two-pass template walks the same row twice
        em.w(f"int {jn} = 0;")
        for cty, nm, _ in loads:
            em.w(f"{cty} {nm}_n = 0;")
        em.w(f"if (kb{gi} < ke{gi}) {{")
        em.w(f"  {jn} = c.p_{canon}__indices[kb{gi}];")
        for _, nm, tx in loads:
            em.w(f"  {nm}_n = {at(tx, f'kb{gi}')};")
        em.w("}")
        em.w(f"for (int kk = kb{gi}; kk < ke{gi}; ++kk) {{")
        em.ind += 1
        em.w(f"const int j = {jn};")
        for cty, nm, _ in loads:
            em.w(f"const {cty} {nm} = {nm}_n;")
        em.w(f"if (kk + 1 < ke{gi}) {{")
        em.w(f"  {jn} = c.p_{canon}__indices[kk + 1];")
        for _, nm, tx in loads:
            em.w(f"  {nm}_n = {at(tx, '(kk + 1)')};")
        em.w("}")
        em.w("(void)j;")

    @staticmethod
    def close_walk(em: _Emitter):
        em.ind -= 1
        em.w("}")
        em.w("}")

    def emit_exprs(self, em: _Emitter, targets: List[Tuple[str, sp.Expr]]):
        """targets: (statement prefix e.g. 'S0 +=' or 'const double v3 =', expr)."""
        exprs = [self.canon(e) for _, e in targets]
        prefix = f"x{self.cse_count}_"
        self.cse_count += 1
        repl, red = sp.cse(exprs, symbols=sp.numbered_symbols(prefix, real=True), order="none")
        pr = _CPrinter(self.names)
        # sincos fusion
        by_arg: Dict[sp.Expr, Dict[str, sp.Symbol]] = {}
        for sym, sub in repl:
            if isinstance(sub, (sp.sin, sp.cos)):
                by_arg.setdefault(sub.args[0], {})["s" if isinstance(sub, sp.sin) else "c"] = sym
        fused: Set[sp.Symbol] = set()
        for sym, sub in repl:
            if sym in fused:
                continue
            if isinstance(sub, (sp.sin, sp.cos)):
                pair = by_arg.get(sub.args[0], {})
                if "s" in pair and "c" in pair:
                    # emit at the position of whichever comes first; both args are available
                    em.w(f"double {pair['s'].name}, {pair['c'].name};")
                    em.w(f"sincos({pr.doprint(sub.args[0])}, &{pair['s'].name}, &{pair['c'].name});")
                    fused.add(pair["s"])
                    fused.add(pair["c"])
                    continue
            em.w(f"const double {sym.name} = {pr.doprint(sub)};")
        for (stmt, _), e in zip(targets, red):
            em.w(f"{stmt} {pr.doprint(e)};")

    def used_loads(self, exprs: Sequence[sp.Expr]) -> Set[sp.Symbol]:
        out: Set[sp.Symbol] = set()
        for e in exprs:
            out |= {s for s in self.canon(e).free_symbols if s in self.eq.loads}
        return out

    def walk_header(self, em: _Emitter, gi: int, group: List[int]):
        w0 = self.eq.walkers[group[0]]
        canon = self.ir.params[w0.matrix + "__indptr"].pattern_of or w0.matrix
        em.w(f"const int r{gi} = {self.idx_text(w0.row)};")
        em.w(f"const int kb{gi} = c.p_{_ident(canon)}__indptr[r{gi}], ke{gi} = c.p_{_ident(canon)}__indptr[r{gi} + 1];")
        return canon


class _Tables:
    """int32 tables printed as static arrays."""

    def __init__(self, trig_ok: bool = True):
        self.items: List[Tuple[str, np.ndarray]] = []
        # variables whose sin / cos are tabulated per scenario (trig hoisting): base -> (table offset, length)
        self.trig: Dict[str, Tuple[int, int]] = {}
        self.trig_ok = trig_ok

    def trig_table(self, ir: ModelIR, base: str) -> None:
        if base not in self.trig:
            off = sum(ln for _, ln in self.trig.values())
            self.trig[base] = (off, int(ir.var_off[base][1]))

    def add(self, name: str, arr: np.ndarray) -> str:
        self.items.append((name, np.asarray(arr, dtype=np.int32)))
        return name

    def text(self) -> str:
        out = []
        for name, arr in self.items:
            body = ",".join(map(str, arr.tolist())) if arr.size else "0"
            out.append(f"SZB_TAB int {name}[{max(arr.size, 1)}] = {{{body}}};")
        return "\n".join(out)


def _gen_F(ir: ModelIR, k: int, eq: EqTemplate, jt: JacTemplate, tabs: _Tables) -> str:
    g = _RowGen(ir, eq, tabs)
    g.set_groups(jt.groups)
    em = _Emitter()
    wexpr = [g.hoist(w.expr) for w in eq.walkers]
    outer_syms = g.used_loads([eq.expr] + wexpr) | set(g.trig_outer)
    g.emit_loads(em, outer_syms, inner=False)
    g.emit_outer_trig(em)
    for gi, group in enumerate(jt.groups):
        canon = g.walk_header(em, gi, group)
        for wi in group:
            em.w(f"double {g.names[eq.walkers[wi].sym]} = 0.0;")
        ex = [wexpr[wi] for wi in group]
        g.open_walk(em, gi, _ident(canon), g.used_loads(ex), g.trig_loads(ex), pipelined=False)
        g.emit_exprs(em, [(f"{g.names[eq.walkers[wi].sym]} +=", wexpr[wi]) for wi in group])
        g.close_walk(em)
    g.emit_exprs(em, [("return", eq.expr)])
    head = f"SZB_HD static inline double szbm_F{k}(const int i, const szbm_ctx& c) {{  // {eq.name}\n"
    return head + em.text() + "\n}\n"


def _gen_J(ir: ModelIR, k: int, jt: JacTemplate, tabs: _Tables) -> str:
    eq = jt.eq
    # no trig hoisting in the Jacobian rows: measured on B200, the J kernel is bound by the latency of its
    # scattered CSC stores, not by sincos, and the two extra table reads per position cost 9 %
    g = _RowGen(ir, eq, tabs, hoist_trig=False)
    g.set_groups(jt.groups)
    for _, a, _ in jt.accs:
        g.names[a] = a.name.lstrip("_")
    em = _Emitter()
    if not jt.outer and not jt.inner:
        return f"SZB_HD static inline void szbm_J{k}(const int, const szbm_ctx&, double* __restrict__) {{}}  // {eq.name}\n"
    wexpr = [g.hoist(w.expr) for w in eq.walkers]                  # walk bodies with the trig hoisted
    terms = [g.hoist(e.term) for e in jt.inner]
    accs_h = [(wi, a, g.hoist(x)) for wi, a, x in jt.accs]
    all_exprs = [e.expr for e in jt.outer] + [e.coef for e in jt.inner] + terms + [x for _, _, x in accs_h]
    need_syms = set().union(*[e.free_symbols for e in all_exprs]) if all_exprs else set()
    walkers_needed = [wi for wi, w in enumerate(eq.walkers) if w.sym in need_syms]
    all_exprs += [wexpr[wi] for wi in walkers_needed]
    g.emit_loads(em, g.used_loads(all_exprs) | set(g.trig_outer), inner=False)
    g.emit_outer_trig(em)
    otab = [tabs.add(f"szbm_ot{k}_{ei}", sl) for ei, sl in enumerate(jt.outer_slots)]
    # Inner entries that hit the same CSC slot at every position of the same walk (e.g. the G-walk and the
    # B-walk terms of dP/dVa) are summed in registers and emitted ONCE: emitted separately the second one is
    # a global read-modify-write of the value the same thread stored an instruction earlier (an L2 round
    # trip on the dependent chain per term).  leader[ei] = the entry whose slot table the sum goes through.
    group_of_w = {wi: gi for gi, g in enumerate(jt.groups) for wi in g}
    leader: Dict[int, int] = {}
    seen: Dict[Tuple, int] = {}
    for ei, e in enumerate(jt.inner):
        sl = np.asarray(jt.inner_slots[ei])
        key = (group_of_w[e.walker], np.where(sl >= 0, sl, ~sl).tobytes(),
               np.asarray(jt.inner_off[e.walker]).tobytes())
        leader[ei] = seen.setdefault(key, ei)
    itab = {ei: tabs.add(f"szbm_it{k}_{ei}", jt.inner_slots[ei]) for ei in sorted(set(leader.values()))}
    woff = {wi: tabs.add(f"szbm_wo{k}_{wi}", off) for wi, off in jt.inner_off.items()
            if any(e.walker == wi and leader[ei] == ei for ei, e in enumerate(jt.inner))}

    def sums_loop(gi, group, do_sums: bool, do_emit: bool):
        canon_ = self_canon[gi]
        accs_here = [(wi, a, x) for wi, a, x in accs_h if wi in group]
        s_here = [wi for wi in group if wi in walkers_needed]
        ents = [(ei, e) for ei, e in enumerate(jt.inner) if e.walker in group]
        if not ((do_sums and (accs_here or s_here)) or (do_emit and ents)):
            return
        targets: List[Tuple[str, sp.Expr]] = []
        if do_sums:
            targets += [(f"{g.names[eq.walkers[wi].sym]} +=", wexpr[wi]) for wi in s_here]
            targets += [(f"{g.names[a]} +=", x) for _, a, x in accs_here]
        slot_loads: List[Tuple[str, str, str]] = []
        if do_emit:
            for ei, e in ents:
                targets.append((f"const double v{ei} =", sp.Symbol(f"cf{ei}", real=True) * terms[ei]))
                if leader[ei] == ei:                   # the CSC slot of this position, fetched an iteration ahead
                    slot_loads.append(("int", f"sl{ei}", f"{itab[ei]}[{woff[e.walker]}[i] + (kk - kb{gi})]"))
        g.open_walk(em, gi, _ident(canon_), g.used_loads([x for _, x in targets]),
                    g.trig_loads([x for _, x in targets]) + slot_loads)
        g.emit_exprs(em, targets)
        if do_emit:
            for ei, e in ents:
                if leader[ei] != ei:
                    continue
                total = " + ".join(f"v{ej}" for ej, _ in ents if leader[ej] == ei)
                em.w(f"szb_emit(JV, sl{ei}, {total});")
        g.close_walk(em)

    self_canon = {}
    for gi, group in enumerate(jt.groups):
        self_canon[gi] = g.walk_header(em, gi, group)
        for wi in group:
            if wi in walkers_needed:
                em.w(f"double {g.names[eq.walkers[wi].sym]} = 0.0;")
        for wi, a, _ in jt.accs:
            if wi in group:
                em.w(f"double {g.names[a]} = 0.0;")
    coef_targets = [(f"const double cf{ei} =", e.coef) for ei, e in enumerate(jt.inner)]
    if jt.two_pass:
        for gi, group in enumerate(jt.groups):
            sums_loop(gi, group, True, False)
        if coef_targets:
            g.emit_exprs(em, coef_targets)
        for gi, group in enumerate(jt.groups):
            sums_loop(gi, group, False, True)
    else:
        if coef_targets:
            g.emit_exprs(em, coef_targets)
        for gi, group in enumerate(jt.groups):
            sums_loop(gi, group, True, True)
    if jt.outer:
        g.emit_exprs(em, [(f"const double o{ei} =", e.expr) for ei, e in enumerate(jt.outer)])
        for ei, e in enumerate(jt.outer):
            em.w(f"szb_emit(JV, {otab[ei]}[i], o{ei});")
    head = (f"SZB_HD static inline void szbm_J{k}(const int i, const szbm_ctx& c, double* __restrict__ JV) {{"
            f"  // {eq.name}\n")
    return head + em.text() + "\n}\n"


_PRELUDE = r'''// Generated by solverz_b200.code_printer.cuda -- do not edit.
// model: %(name)s   kind: %(kind)s   n = %(n)d   nnz(J) = %(nnz)d   equations: %(neq)d
#include <stdint.h>
#include <math.h>
#include "szb_model.h"

#ifdef SZBM_HOST_CHECK
#define SZB_HD
#define SZB_TAB static const
#else
#include <cuda_runtime.h>
#define SZB_HD __host__ __device__
#define SZB_TAB static __device__ const
#endif

SZB_HD static inline double szb_sign(double x) { return (double)((x > 0.0) - (x < 0.0)); }
SZB_HD static inline double szb_heaviside(double x) { return x >= 0.0 ? 1.0 : 0.0; }
SZB_HD static inline double szb_saturation(double v, double lo, double hi) { return fmin(fmax(v, lo), hi); }
SZB_HD static inline double szb_in(double v, double lo, double hi) { return (v >= lo && v <= hi) ? 1.0 : 0.0; }
SZB_HD static inline double szb_gt(double a, double b) { return a > b ? 1.0 : 0.0; }
SZB_HD static inline double szb_lt(double a, double b) { return a < b ? 1.0 : 0.0; }
SZB_HD static inline double szb_and(double a, double b) { return (a != 0.0 && b != 0.0) ? 1.0 : 0.0; }
SZB_HD static inline double szb_or(double a, double b) { return (a != 0.0 || b != 0.0) ? 1.0 : 0.0; }
SZB_HD static inline double szb_not(double a) { return a != 0.0 ? 0.0 : 1.0; }
SZB_HD static inline double szb_min(double a, double b) { return a <= b ? a : b; }
// TimeSeriesParam.get_v_t (Solverz/equation/param.py:187-212): linear for t < t_end, else last value
SZB_HD static inline double szb_interp(const double* ts, const double* vs, int nk, double t) {
  if (!(t < ts[nk - 1])) return vs[nk - 1];
  if (t <= ts[0]) return vs[0];
  int lo = 0, hi = nk - 1;
  while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (ts[mid] <= t) lo = mid; else hi = mid; }
  return vs[lo] + (vs[hi] - vs[lo]) * ((t - ts[lo]) / (ts[hi] - ts[lo]));
}
// first emission of a CSC slot stores, later ones (slot index bit-complemented) accumulate
SZB_HD static inline void szb_emit(double* __restrict__ JV, const int slot, const double v) {
  if (slot >= 0) JV[slot] = v; else JV[~slot] += v;
}
'''

_KERNELS = r'''
#ifndef SZBM_HOST_CHECK
struct szbm_kargs {
  int B; double t; const double* t_b; const double* Y; double* F; double* JV; double* fnorm; const int* active;
%(karg_fields)s
};

__device__ __forceinline__ szbm_ctx szbm_make_ctx(const szbm_kargs& a, const int b) {
  szbm_ctx c;
  c.t = a.t_b ? a.t_b[b] : a.t;
  c.Y = a.Y + (size_t)b * SZBM_N;
%(ctx_fill)s
  return c;
}

__device__ __forceinline__ double szbm_nanmax(double a, double b) { return (a > b || a != a) ? a : b; }

__global__ void __launch_bounds__(SZBM_THREADS) szbm_F_kernel(const szbm_kargs a) {
  __shared__ double red[SZBM_THREADS / 32];
#if SZBM_TRIG_N > 0
  __shared__ double szbm_ts[SZBM_TRIG_N], szbm_tc[SZBM_TRIG_N];
#endif
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    if (a.active && !a.active[b]) continue;
    szbm_ctx c = szbm_make_ctx(a, b);
#if SZBM_TRIG_N > 0
    __syncthreads();                      // the rows of the previous scenario are done with the tables
    szbm_fill_trig(c, szbm_ts, szbm_tc, threadIdx.x, SZBM_THREADS);
    __syncthreads();
#endif
    double* __restrict__ F = a.F + (size_t)b * SZBM_N;
    double m = 0.0;
    for (int row = threadIdx.x; row < SZBM_N; row += SZBM_THREADS) {
      const double v = szbm_row_F(row, c);
      F[row] = v;
      m = szbm_nanmax(m, fabs(v));
    }
    for (int o = 16; o > 0; o >>= 1) m = szbm_nanmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < SZBM_THREADS / 32; ++w) m = szbm_nanmax(m, red[w]);
      if (a.fnorm) a.fnorm[b] = m;
    }
  }
}

__global__ void __launch_bounds__(SZBM_THREADS) szbm_J_kernel(const szbm_kargs a) {
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    if (a.active && !a.active[b]) continue;
    const szbm_ctx c = szbm_make_ctx(a, b);     // the Jacobian rows read no trig tables (see _gen_J)
    double* __restrict__ JV = a.JV + (size_t)b * SZBM_NNZ;
    for (int row = threadIdx.x; row < SZBM_N; row += SZBM_THREADS) szbm_row_J(row, c, JV);
  }
}

static int szbm_launch(const szbm_args* a, void* stream, int which) {
  if (!a || a->B <= 0) return (int)cudaErrorInvalidValue;
  szbm_kargs k;
  k.B = a->B; k.t = a->t; k.t_b = a->t_b; k.Y = a->Y; k.F = a->F; k.JV = a->JV; k.fnorm = a->fnorm; k.active = a->active;
%(karg_fill)s
  int grid = a->B < 65535 * 16 ? a->B : 65535 * 16;
  if (which == 0) szbm_F_kernel<<<grid, SZBM_THREADS, 0, (cudaStream_t)stream>>>(k);
  else szbm_J_kernel<<<grid, SZBM_THREADS, 0, (cudaStream_t)stream>>>(k);
  return (int)cudaGetLastError();
}
#endif

extern "C" {
int szbm_abi_version(void) { return SZBM_ABI_VERSION; }
int szbm_n(void) { return SZBM_N; }
int szbm_nnz_j(void) { return SZBM_NNZ; }
int szbm_n_params(void) { return SZBM_NPARAMS; }
const szbm_param_info* szbm_params(void) { return szbm_param_table; }
#ifndef SZBM_HOST_CHECK
int szbm_launch_F(const szbm_args* a, void* stream) { return szbm_launch(a, stream, 0); }
int szbm_launch_J(const szbm_args* a, void* stream) { return szbm_launch(a, stream, 1); }
#else
// host check build: same row functions, serial loops; used by tests only
SZBM_API void szbm_host_F(int B, double t, const double* Y, const void* const* params, double* F) {
  for (int b = 0; b < B; ++b) {
    szbm_ctx c; c.t = t; c.Y = Y + (size_t)b * SZBM_N;
%(host_ctx_fill)s
    static double ts_[SZBM_TRIG_N > 0 ? SZBM_TRIG_N : 1], tc_[SZBM_TRIG_N > 0 ? SZBM_TRIG_N : 1];
    szbm_fill_trig(c, ts_, tc_, 0, 1);
    for (int row = 0; row < SZBM_N; ++row) F[(size_t)b * SZBM_N + row] = szbm_row_F(row, c);
  }
}
SZBM_API void szbm_host_J(int B, double t, const double* Y, const void* const* params, double* JV) {
  for (int b = 0; b < B; ++b) {
    szbm_ctx c; c.t = t; c.Y = Y + (size_t)b * SZBM_N;
%(host_ctx_fill)s
    static double ts_[SZBM_TRIG_N > 0 ? SZBM_TRIG_N : 1], tc_[SZBM_TRIG_N > 0 ? SZBM_TRIG_N : 1];
    szbm_fill_trig(c, ts_, tc_, 0, 1);
    for (int row = 0; row < SZBM_N; ++row) szbm_row_J(row, c, JV + (size_t)b * SZBM_NNZ);
  }
}
#endif
}
'''


def print_model(ir: ModelIR, pat: JacPattern, name: str = "model", threads: int = 256) -> str:
    """Return the CUDA source of the model library."""
    # Trig hoisting needs 16 bytes of static shared memory per tabulated variable entry (48 kB limit):
    # generate with it, and again without it when the tables would not fit.
    for trig_ok in (True, False):
        tabs = _Tables(trig_ok)
        _CPrinter.snippets = {}
        f_funcs, j_funcs = [], []
        for k, (eq, jt) in enumerate(zip(ir.eqs, pat.templates)):
            f_funcs.append(_gen_F(ir, k, eq, jt, tabs))
            j_funcs.append(_gen_J(ir, k, jt, tabs))
        trig_n = sum(ln for _, ln in tabs.trig.values())
        if trig_n <= 3072:
            break
    pnames = list(ir.params)
    ctx_fields = ["  double t;", "  const double* Y;"]
    for base in tabs.trig:
        ctx_fields.append(f"  const double* ts_{_ident(base)};")
        ctx_fields.append(f"  const double* tc_{_ident(base)};")
    fill = ["SZB_HD static inline void szbm_fill_trig(szbm_ctx& c, double* ts, double* tc, const int tid, const int nthr) {",
            "  (void)ts; (void)tc; (void)tid; (void)nthr; (void)c;"]
    for base, (toff, ln) in tabs.trig.items():
        voff = ir.var_off[base][0]
        fill.append(f"  for (int k = tid; k < {ln}; k += nthr) sincos(c.Y[{voff} + k], ts + {toff} + k, tc + {toff} + k);")
        fill.append(f"  c.ts_{_ident(base)} = ts + {toff}; c.tc_{_ident(base)} = tc + {toff};")
    fill.append("}")
    karg_fields, ctx_fill, karg_fill, host_fill, ptable = [], [], [], [], []
    for pi, pn in enumerate(pnames):
        p = ir.params[pn]
        cty = "double" if p.dtype == "f64" else "int"
        ident = _ident(pn)
        ctx_fields.append(f"  const {cty}* p_{ident};")
        karg_fields.append(f"  const {cty}* p_{ident};")
        stride = f" + (size_t)b * {p.length}" if p.per_scenario else ""
        ctx_fill.append(f"  c.p_{ident} = a.p_{ident}{stride};")
        karg_fill.append(f"  k.p_{ident} = (const {cty}*)a->params[{pi}];")
        host_fill.append(f"    c.p_{ident} = (const {cty}*)params[{pi}]{stride};")
        ptable.append(f'  {{"{pn}", {0 if p.dtype == "f64" else 1}, {1 if p.per_scenario else 0}, {p.length}}},')
    # row dispatch
    disp_F, disp_J = [], []
    for k, eq in enumerate(ir.eqs):
        cond = f"if (row < {eq.row0 + eq.n_rows})" if k < len(ir.eqs) - 1 else ""
        els = "else " if k > 0 else ""
        disp_F.append(f"  {els}{cond} return szbm_F{k}(row - {eq.row0}, c);")
        disp_J.append(f"  {els}{cond} szbm_J{k}(row - {eq.row0}, c, JV);")
    src = _PRELUDE % dict(name=name, kind=ir.kind, n=ir.n, nnz=pat.nnz, neq=len(ir.eqs))
    src += f"\n#define SZBM_N {ir.n}\n#define SZBM_NNZ {pat.nnz}\n#define SZBM_NPARAMS {len(pnames)}\n"
    src += f"#define SZBM_THREADS {threads}\n#define SZBM_TRIG_N {trig_n}\n\n"
    src += "struct szbm_ctx {\n" + "\n".join(ctx_fields) + "\n};\n\n"
    src += "static const szbm_param_info szbm_param_table[] = {\n" + "\n".join(ptable) + \
           ("\n" if ptable else "") + '  {0, 0, 0, 0}\n};\n\n'
    src += "\n".join(fill) + "\n\n" + tabs.text() + "\n\n"
    if _CPrinter.snippets:
        src += "// device code of custom functions (_cuda_device_code)\n" + "\n".join(_CPrinter.snippets.values()) + "\n\n"
    src += "\n".join(f_funcs) + "\n" + "\n".join(j_funcs) + "\n"
    src += "SZB_HD static inline double szbm_row_F(const int row, const szbm_ctx& c) {\n" + "\n".join(disp_F) + "\n}\n"
    src += "SZB_HD static inline void szbm_row_J(const int row, const szbm_ctx& c, double* __restrict__ JV) {\n" + \
           "\n".join(disp_J) + "\n}\n"
    src += _KERNELS % dict(karg_fields="\n".join(karg_fields), ctx_fill="\n".join(ctx_fill),
                           karg_fill="\n".join(karg_fill), host_ctx_fill="\n".join(host_fill))
    return src


def source_hash(src: str) -> str:
    return hashlib.sha256(src.encode()).hexdigest()[:16]
