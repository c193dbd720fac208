"""Batched fixed-step implicit one-step methods with the reference's surface:
``backward_euler(dae, tspan, y0, opt)`` (Solverz/solvers/daesolver/beuler.py:4-91) and
``implicit_trapezoid(dae, tspan, y0, opt)`` (Solverz/solvers/daesolver/trapezoidal.py:4-92).

The reference wraps every step into an ``nAE`` closure and hands it to ``nr_method``; here every step is
one batched ``szb_newton`` in its implicit one-step mode (``szb_set_implicit``): the residual
``M (y - y0) - c (F(t0 + dt, y) + kappa F(t0, y0))`` is formed on the device from the model's F, the matrix
``M - c J`` inside the factorisation kernel, warm start from the previous node.  The counters follow the
reference line by line, including what it does not count (beuler.py:74-75 adds ``nstep`` to both
``ndecomp`` and ``nfeval`` and nothing else).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ..num_api import lib as L
from ..num_api.model import CudaModel
from ..variables import TimeVars, Vars
from .option import Opt
from .rodas import dae_ic, _profiled
from .solution import daesol
from .stats import Stats

__all__ = ["backward_euler", "implicit_trapezoid"]


def _run(dae: CudaModel, tspan, y0, opt: Opt, mode: int, scheme: str):
    import torch
    if not isinstance(dae, CudaModel) or dae.kind != "DAE":
        raise TypeError(f"{scheme} of solverz_b200 needs a DAE model rendered with lang='cuda'")
    stats = Stats(scheme=scheme)
    if opt is None:
        opt = Opt(stats=True)
    dt = opt.step_size
    tspan = np.array(tspan, dtype=float)
    t_init, t_end = float(tspan[0]), float(tspan[-1])
    vars_in = y0 if isinstance(y0, Vars) else None
    y_arr = y0.array if isinstance(y0, Vars) else y0
    on_device = isinstance(y_arr, torch.Tensor)
    if not on_device:
        y_arr = np.asarray(y_arr, dtype=float)
    single = y_arr.ndim == 1
    B = 1 if single else int(y_arr.shape[0])
    s = dae.session(B)
    s.set_params(dae.p)
    s.set_y(y_arr)
    n = dae.n
    rc = np.full(n, -1, dtype=np.int32)
    rc[np.asarray(dae.mass[0], dtype=np.int64)] = np.asarray(dae.mass[1], dtype=np.int32)
    L.check(s.lib.szb_set_mass_rows(s.h, rc.ctypes.data_as(C.POINTER(C.c_int32)), n), s.h, "szb_set_mass_rows")
    dae_ic(s, t_init, opt.rtol)                                # check and modify initial values
    y_prev = s.Y.clone()
    f_prev = torch.zeros_like(y_prev) if mode == 2 else None
    T = [t_init]
    Ys = [s.Y.clone()]
    tt = t0 = t_init
    cnt = {k: torch.zeros(B, dtype=torch.int64, device=s.Y.device) for k in ("nstep", "nfeval")}
    ok = torch.ones(B, dtype=torch.bool, device=s.Y.device)
    # beuler.py:73 calls nr_method with a default Opt (ite_tol 1e-5); trapezoidal.py:74 passes opt.ite_tol
    tol = Opt().ite_tol if mode == 1 else opt.ite_tol
    nt = 0
    try:
        while t_end - tt > abs(dt) / 10:
            y_prev.copy_(s.Y)
            if mode == 2:
                f_prev.copy_(s.eval_F(t0))                     # F0 = dae.F(t0, y0, p), trapezoidal.py:67
            c = dt if mode == 1 else dt / 2
            L.check(s.lib.szb_set_implicit(s.h, mode, float(c), C.c_void_p(y_prev.data_ptr()),
                                           C.c_void_p(f_prev.data_ptr()) if mode == 2 else None), s.h,
                    "szb_set_implicit")
            s.set_time(t0 + dt)
            iters, flags = s.newton(tol, Opt().max_it)
            cnt["nstep"] += iters.to(torch.int64)
            ok &= flags == 0
            tt = tt + dt
            nt += 1
            T.append(tt)
            Ys.append(s.Y.clone())
            t0 = tt
    finally:
        L.check(s.lib.szb_set_implicit(s.h, 0, 0.0, None, None), s.h, "szb_set_implicit")
        s.set_time(0.0)
    it = cnt["nstep"].cpu().numpy()
    stats.nstep = nt
    if mode == 1:                                              # beuler.py:74-75
        ndecomp, nfeval, nJ, nsolve = it, it, np.zeros_like(it), np.zeros_like(it)
    else:                                                      # trapezoidal.py:72-79: + F0 per step, nr: nfeval = nstep + 1
        ndecomp, nfeval, nJ, nsolve = it, it + 2 * nt, it, it
    pick = (lambda a: int(a[0])) if single else (lambda a: a.copy())
    stats.ndecomp, stats.nfeval, stats.nJeval, stats.nsolve = pick(ndecomp), pick(nfeval), pick(nJ), pick(nsolve)
    stats.succeed = bool(ok[0]) if single else ok.cpu().numpy()
    Y = torch.stack(Ys, dim=1)                                 # [B, nt+1, n]
    T = np.array(T)
    if single:
        Y = Y[0].cpu().numpy()
        if vars_in is not None:
            Y = TimeVars(vars_in.a, Y)
    elif not on_device:
        Y = Y.cpu().numpy()
    return daesol(T, Y, stats=stats)


@_profiled
def backward_euler(dae: CudaModel, tspan, y0, opt: Opt = None):
    return _run(dae, tspan, y0, opt, 1, 'Backward Euler')


@_profiled
def implicit_trapezoid(dae: CudaModel, tspan, y0, opt: Opt = None):
    return _run(dae, tspan, y0, opt, 2, 'Trapezoidal')
