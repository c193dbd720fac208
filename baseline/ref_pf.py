"""Reference arm of bench.py: the UNMODIFIED reference (private copy under baseline/_ref, see make_ref.sh)
solving the bench workload on the host cores through its own public API and stock code path --
``module_printer(spf, y0, name, jit=True).render()`` -> numba ``@njit`` module, then per scenario
``mdl.p['Pinj'] = ...; mdl.p['Qinj'] = ...; nr_method(mdl, y0, Opt(ite_tol=1e-8))``.

Nothing of this repo's product is on that path: only the synthetic grid / scenario generators
(``solverz_b200.synthetic``: numpy) are shared so that both arms solve the same numbers.
"""
from __future__ import annotations

import importlib
import os
import sys
import time
from pathlib import Path

import numpy as np
import scipy.sparse as sps

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref"
MODELS = HERE / "_ref" / "_rendered"


def available() -> str:
    """'' when the reference copy and numba are importable, else the reason."""
    if not (REF / "Solverz" / "__init__.py").exists():
        return "baseline/_ref is missing (baseline/make_ref.sh needs /root/reference)"
    try:
        import numba  # noqa: F401
    except Exception as e:  # pragma: no cover
        return f"numba not importable: {e}"
    return ""


def _paths():
    for p in (str(REF), str(HERE.parent)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.setdefault("NUMBA_CACHE_DIR", str(REF / "_numba_cache"))


def render(nx: int) -> str:
    """Build the polar power flow of the nx x nx lattice with the reference's DSL (the form SURVEY.md 8d
    probed: P_eqn over PVPQ, Q_eqn over PQ, G and B walked in separate Sums) and render its numba module."""
    _paths()
    name = f"ref_pf_{nx}x{nx}"
    if (MODELS / name / "num_func.py").exists():
        return name
    from Solverz import Eqn, LoopEqn, Model, Param, Set, Sum, Var, cos, module_printer, sin
    from solverz_b200.synthetic import lattice_grid
    nb, Y = lattice_grid(nx, nx)
    G = sps.csc_array(Y.real)
    B = sps.csc_array(Y.imag)
    npv = max(1, int(0.1 * nb))
    pv = list(range(1, 1 + npv))
    pq = list(range(1 + npv, nb))
    rng = np.random.default_rng(1)
    Pb = rng.uniform(-.05, .05, nb)
    Qb = rng.uniform(-.02, .02, nb)
    m = Model()
    m.Va = Var('Va', np.zeros(nb))
    m.Vm = Var('Vm', np.ones(nb))
    m.G = Param('G', G, dim=2, sparse=True)
    m.B = Param('B', B, dim=2, sparse=True)
    m.Pinj = Param('Pinj', Pb)
    m.Qinj = Param('Qinj', Qb)
    m.Bus = Set('Bus', nb)
    m.PVPQ = Set('PVPQ', np.array(pv + pq))
    m.PQ = Set('PQ', np.array(pq))
    i = m.PVPQ.idx('i')
    j = m.Bus.idx('j')
    m.P_eqn = LoopEqn('P_eqn', outer_index=i,
                      body=m.Vm[i] * Sum(m.Vm[j] * m.G[i, j] * cos(m.Va[i] - m.Va[j]), j)
                      + m.Vm[i] * Sum(m.Vm[j] * m.B[i, j] * sin(m.Va[i] - m.Va[j]), j) - m.Pinj[i], model=m)
    k = m.PQ.idx('k')
    m.Q_eqn = LoopEqn('Q_eqn', outer_index=k,
                      body=m.Vm[k] * Sum(m.Vm[j] * m.G[k, j] * sin(m.Va[k] - m.Va[j]), j)
                      - m.Vm[k] * Sum(m.Vm[j] * m.B[k, j] * cos(m.Va[k] - m.Va[j]), j) - m.Qinj[k], model=m)
    m.ref_eqn = Eqn('Va_ref', m.Va[0])
    m.vm_eqn = Eqn('Vm_fix', m.Vm[0:1 + npv] - 1)
    spf, y0 = m.create_instance()
    MODELS.mkdir(parents=True, exist_ok=True)
    module_printer(spf, y0, name, directory=str(MODELS), jit=True).render()
    return name


_STATE = {}


def _load(nx: int):
    if nx not in _STATE:
        _paths()
        os.environ["OMP_NUM_THREADS"] = os.environ["MKL_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
        if str(MODELS) not in sys.path:
            sys.path.insert(0, str(MODELS))
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):         # the module prints its compile time: not on bench's stdout
            mod = importlib.import_module(f"ref_pf_{nx}x{nx}")   # importing JIT-compiles F_ / J_ once (cached on disk)
        from Solverz import Opt, nr_method
        from solverz_b200.synthetic import lattice_grid
        nb, _ = lattice_grid(nx, nx)
        rng = np.random.default_rng(1)
        info = dict(nb=nb, Pbase=rng.uniform(-.05, .05, nb), Qbase=rng.uniform(-.02, .02, nb))
        _STATE[nx] = (mod.mdl, mod.y, info, Opt, nr_method)
    return _STATE[nx]


def solve_block(args):
    """Worker: scenarios [first, first+count) by the reference's nr_method; returns (Newton iterations,
    solutions or None)."""
    nx, first, count, tol, keep = args
    mdl, y0, info, Opt, nr_method = _load(nx)
    from solverz_b200.synthetic import pf_scenarios
    P, Q = pf_scenarios(info, count, first)
    its = 0
    ys = []
    for b in range(count):
        mdl.p["Pinj"] = P[b]
        mdl.p["Qinj"] = Q[b]
        s = nr_method(mdl, y0, Opt(ite_tol=tol))
        its += int(s.stats.nstep)
        if keep:
            ys.append((s.y.array.copy(), int(s.stats.nstep)))
    return its, ys


def timed_run(nx: int, sample: int, cores: int, steps: int, warmup: int, tol: float):
    """(iterations/s, iterations per step, seconds per step, scenarios per step) on `cores` worker processes."""
    import multiprocessing as mp
    render(nx)
    per = max(1, sample // cores)
    blocks = [(nx, i * per, per, tol, False) for i in range(cores)]
    ctx = mp.get_context("spawn")          # numba + fork do not mix reliably
    # the initializer imports the rendered module in EVERY worker (numba compile / cache load) before any timing
    with ctx.Pool(cores, initializer=_load, initargs=(nx,)) as pool:
        for _ in range(max(warmup, 1)):
            pool.map(solve_block, [(nx, 0, 1, tol, False)] * cores)
        t0 = time.perf_counter()
        total = 0
        for _ in range(steps):
            total += sum(r[0] for r in pool.map(solve_block, blocks))
        dt = time.perf_counter() - t0
    return total / dt, total // max(steps, 1), dt / max(steps, 1), per * cores
