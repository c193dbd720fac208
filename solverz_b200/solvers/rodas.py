"""Batched Rodas with the reference's ``Rodas(dae, tspan, y0, opt)`` surface.

Reference: Solverz/solvers/daesolver/rodas/rodas.py:7-344 and the ``dae_io_parser`` decorator
(solvers/parser.py:65-121).  The step loop runs in ``szb_rodas`` (csrc/szb_rodas.cu); this module
does the consistent initialisation (``DaeIc``, daeic.py:24-68) and moves data.

* ``y0`` 1-D / ``Vars``: one scenario, returns ``daesol(T, Y, stats)`` like the reference -- every
  accepted step for a two-point ``tspan``, dense output at the given nodes otherwise.
* ``y0`` ``[B, n]``: a batch; ``T`` is ``tspan`` (dense output) or ``[t0, t_reached]``, ``Y`` has shape
  ``[B, nt, n]`` and the ``Stats`` counters are int arrays of length ``B``.
* ``opt.dfdt`` (backend knob, default ``'module'``): ``'module'`` reproduces the reference's rendered
  numba module, where dF/dt evaluates to 0; ``'fd'`` uses the forward difference
  of rodas.py:347-352 as the inline ``made_numerical`` path does.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ..num_api import lib as L
from ..num_api.model import CudaModel
from ..variables import TimeVars, Vars
from .option import Opt
from .rodas_param import Rodas_param
from .solution import daesol
from .stats import Stats


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def dae_ic(sess, t0: float, rtol: float):
    """Batched restatement of DaeIc (daeic.py:24-68): Newton on the algebraic sub-system
    J[AlgEqn, AlgVar] with the three-probe halving line search.  The sub-system is solved with the
    full-size matrix  A = M + diag(alg rows) J  whose differential rows are unit rows."""
    import torch
    mdl = sess.model
    n = mdl.n
    de, dv = mdl.mass
    dev = sess.Y.device
    alg_eqn = torch.as_tensor(np.setdiff1d(np.arange(n), de), device=dev)
    alg_var = torch.as_tensor(np.setdiff1d(np.arange(n), dv), device=dev)
    tt = torch.full((sess.B,), float(t0), dtype=torch.float64, device=dev)
    F0 = sess.eval_F(t0)
    need = torch.linalg.norm(F0[:, alg_eqn], dim=1) > 1e-6
    if not bool(need.any()):
        return
    # 0/1 mask per Jacobian slot: 1 on algebraic rows
    rows = mdl.indices
    is_alg_row = np.ones(n)
    is_alg_row[de] = 0.0
    mask = torch.as_tensor(is_alg_row[rows], device=dev)
    y = sess.Y.clone()
    done = ~need

    def solve_alg(yy, rhs_alg):
        sess.Y.copy_(yy)
        sess.eval_J(t0)
        L.check(sess.lib.szb_lu_refactor_masked_rows(sess.h, C.c_void_p(mask.data_ptr()), sess._stream()),
                sess.h, "szb_lu_refactor_masked_rows")
        bad = sess.lu_status() != 0
        if bool(bad.any()):
            # the reference's spsolve raises / warns on a singular J[AlgEqn, AlgVar] (daeic.py:38 -> laesolver.py:11-15)
            raise ValueError(f"DaeIc: the algebraic sub-Jacobian is singular for scenario(s) "
                             f"{torch.nonzero(bad).flatten().tolist()[:8]} (index > 1, or an inconsistent model)")
        sess.RHS.zero_()
        sess.RHS[:, alg_eqn] = rhs_alg
        x = sess.lu_solve(1)
        return x[:, alg_var].clone()

    tiny = float(np.spacing(t0))
    for _ in range(15):
        sess.Y.copy_(y)
        Fn = sess.eval_F(t0).clone()
        dY = solve_alg(y, -Fn[:, alg_eqn])
        ya = y[:, alg_var]
        nz = ya.abs() > tiny
        res = torch.linalg.norm(torch.where(nz, dY / torch.where(nz, ya, torch.ones_like(ya)), torch.zeros_like(ya)), dim=1)
        lam = torch.ones(sess.B, dtype=torch.float64, device=dev)
        settled = done.clone()          # scenarios whose probe loop has ended
        ynew = y.clone()
        resnew = torch.zeros_like(res)
        for _probe in range(3):
            cand = y.clone()
            cand[:, alg_var] = ya + lam[:, None] * dY
            ynew = torch.where(settled[:, None], ynew, cand)
            sess.Y.copy_(ynew)
            Fnew = sess.eval_F(t0).clone()
            ok = torch.linalg.norm(Fnew[:, alg_eqn], dim=1) <= 1e-5 * rtol
            newly = ok & ~settled
            y = torch.where(newly[:, None], ynew, y)
            done = done | newly
            settled = settled | newly
            dYn = solve_alg(ynew, Fnew[:, alg_eqn])
            rn = torch.linalg.norm(torch.where(nz, dYn / torch.where(nz, ya, torch.ones_like(ya)), torch.zeros_like(ya)), dim=1)
            resnew = torch.where(settled, resnew, rn)
            brk = (rn < 0.9 * res) & ~settled
            settled = settled | brk
            lam = torch.where(settled, lam, 0.5 * lam)
        y = torch.where(done[:, None], y, ynew)
        done = done | (resnew <= 1e-3 * rtol)
        if bool(done.all()):
            sess.Y.copy_(y)
            return
    raise ValueError("Need Better y0")


def _finish_callback_events(dae, s, tspan, opt, rp, o, out, y_out, stats_d, single, dense, vars_in, on_device):
    """``Opt(event=callable)``: ``szb_rodas`` one accepted step per call, the callback in between
    (solvers/rodas_event.py); assembles the reference's return value (rodas.py:329-342)."""
    from .rodas_event import rodas_callback_events
    B, n = s.B, dae.n
    o.stop_after_accepted = 1

    def run_call(resume: bool):
        o.resume = 1 if resume else 0
        L.check(s.lib.szb_rodas(s.h, C.byref(o), C.byref(out), s._stream()), s.h, "szb_rodas")

    _, st, loc, T_steps, Y_steps = rodas_callback_events(dae, s, tspan, opt, rp, run_call, single, vars_in)
    st = st.astype(np.int64)
    stats = Stats(opt.scheme)
    for i, k in enumerate(("nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve")):
        setattr(stats, k, int(st[0, i]) if single else st[:, i].copy())
    stats.flag = int(st[0, 6]) if single else st[:, 6].copy()
    failed = (st[:, 6] != 0) & (st[:, 6] != L.SZB_FLAG_EVENT)
    stats.ret = ("failed" if failed[0] else None) if single else np.where(failed, "failed", None)
    stats.succeed = False if single else ~failed
    nev = np.array([len(l.te) for l in loc], dtype=np.int64)
    stats.nevent = int(nev[0]) if single else nev
    Ts, Ys = [], []
    for b in range(B):
        if dense:
            te_last = loc[b].te[-1] if loc[b].terminal else None
            rows = len(tspan) if te_last is None else int(np.searchsorted(tspan, te_last, side='right'))
            if failed[b]:
                rows = min(rows, int(np.searchsorted(tspan, T_steps[b][-1], side='right')))
            Tb = tspan[:rows].copy()
            Yb = y_out[:rows, b].cpu().numpy()
            if te_last is not None and (rows == 0 or Tb[-1] < te_last):     # the event point closes T (rodas.py:291-299)
                Tb = np.append(Tb, te_last)
                Yb = np.vstack([Yb, loc[b].ye[-1][None, :]])
        else:
            Tb, Yb = np.array(T_steps[b]), np.array(Y_steps[b])
        Ts.append(Tb)
        Ys.append(Yb)
    if single:
        T, Y = Ts[0], Ys[0]
        te = np.array(loc[0].te) if loc[0].te else None
        ye = np.array(loc[0].ye) if loc[0].te else None
        ie = np.array(loc[0].ie, dtype=float) if loc[0].te else None
        if vars_in is not None:
            Y = TimeVars(vars_in.a, Y)
            if ye is not None:
                ye = TimeVars(vars_in.a, ye)
        return daesol(T, Y, te, ye, ie, stats=stats)
    # a batch: ragged results as lists (one entry per scenario)
    te = [np.array(l.te) for l in loc]
    ye = [np.array(l.ye) for l in loc]
    ie = [np.array(l.ie, dtype=float) for l in loc]
    return daesol(Ts, Ys, te, ye, ie, stats=stats)


def _profiled(fn):
    """``Opt(profile=True)``: wall time of the whole call, printed like the reference's dae_io_parser does
    (solvers/parser.py:91-96)."""
    import functools
    import time

    @functools.wraps(fn)
    def wrapper(dae, tspan, y0, opt: Opt = None, **kw):
        if opt is not None and getattr(opt, "profile", False):
            import torch
            torch.cuda.synchronize()
            start = time.perf_counter()
            sol = fn(dae, tspan, y0, opt, **kw)
            torch.cuda.synchronize()
            print(f"Time elapsed: {time.perf_counter() - start}s")
            return sol
        return fn(dae, tspan, y0, opt, **kw)
    return wrapper


@_profiled
def Rodas(dae: CudaModel, tspan, y0, opt: Opt = None):
    if opt is None:
        opt = Opt()
    if not isinstance(dae, CudaModel) or dae.kind != "DAE":
        raise TypeError("Rodas of solverz_b200 needs a DAE model rendered with lang='cuda'")
    import torch
    rp = Rodas_param(opt.scheme)
    tspan = np.asarray(tspan, dtype=float)
    t0, tend = float(tspan[0]), float(tspan[-1])
    if t0 > tend:
        raise ValueError(f't0: {t0} > tend: {tend}')
    dense = len(tspan) > 2
    from .rodas_event import ComponentEvents
    dev_events = opt.event if isinstance(opt.event, ComponentEvents) and rp.has_dense_output else None
    callback = opt.event if (opt.event is not None and dev_events is None) else None
    if dense and opt.scheme == "rodas3d":
        # the reference's tableau of rodas3d defines no interpolation coefficients (param.py:171-199: commented out),
        # so its ntrp1 raises there too; rodas3 interpolates with the Hermite form ntrp2 on the device
        raise NotImplementedError("dense output of rodas3d: the scheme has no interpolation coefficients "
                                  "(rodas/param.py:171-199)")
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
    dae_ic(s, t0, opt.rtol)
    if opt.hmax is None:
        opt.hmax = abs(tend - t0)                     # rodas.py:76-77 (the reference mutates opt too)
    dev = s.Y.device
    o = L.RodasOptC()
    o.t0, o.tend = t0, tend
    o.rtol, o.atol, o.f_savety = opt.rtol, opt.atol, opt.f_savety
    o.fac1, o.fac2, o.facmax = opt.fac1, opt.fac2, opt.facmax
    o.hinit = -1.0 if opt.hinit is None else float(opt.hinit)
    o.hmax = float(opt.hmax)
    o.fix_h = 1 if opt.fix_h else 0
    o.dfdt_mode = 1 if getattr(opt, "dfdt", "module") == "fd" else 0
    o.max_steps = int(getattr(opt, "max_steps", 0))
    keep = [np.ascontiguousarray(getattr(rp, k), dtype=np.float64) for k in
            ("alpha", "gammatilde", "a", "g", "b", "bd", "c", "d", "e")]
    o.s, o.pord, o.gamma = rp.s, rp.pord, rp.gamma
    (o.alpha, o.gammatilde, o.a, o.g, o.b, o.bd, o.c, o.d, o.e) = [_dp(k) for k in keep]
    if not rp.has_dense_output:          # no interpolation coefficients: szb_rodas interpolates with ntrp2 (rodas3)
        o.c = o.d = o.e = None
    out = L.RodasOutC()
    t_end = torch.zeros(B, dtype=torch.float64, device=dev)
    stats_d = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    out.t_end, out.stats = t_end.data_ptr(), stats_d.data_ptr()
    y_out = y_steps = t_steps = None
    if dense:
        o.n_out = len(tspan)
        tsp = np.ascontiguousarray(tspan)
        o.t_out = _dp(tsp)
        y_out = torch.zeros((len(tspan), B, n), dtype=torch.float64, device=dev)
        out.y_out = y_out.data_ptr()
        n_rows_all = torch.zeros(B, dtype=torch.int32, device=dev)     # rows of y_out each scenario filled
        out.n_rows = n_rows_all.data_ptr()
    elif single and callback is None:
        cap = int(getattr(opt, "record_cap", 10001))
        o.record_cap = cap
        y_steps = torch.zeros((cap, B, n), dtype=torch.float64, device=dev)
        t_steps = torch.zeros((cap, B), dtype=torch.float64, device=dev)
        out.y_steps, out.t_steps = y_steps.data_ptr(), t_steps.data_ptr()
    ev_count = ev_t = ev_i = ev_y = n_rows = None
    if dev_events is not None:
        ev = dev_events
        if int(ev.index.max()) >= n or int(ev.index.min()) < 0:
            raise IndexError(f"ComponentEvents: component index out of range for n = {n}")
        cap = ev.max_events
        keep += [np.ascontiguousarray(ev.index), np.ascontiguousarray(ev.threshold), np.ascontiguousarray(ev.direction),
                 np.ascontiguousarray(ev.terminal)]
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        o.n_events, o.ev_idx, o.ev_thr, o.ev_dir, o.ev_term = len(ev.index), ip(keep[-4]), _dp(keep[-3]), ip(keep[-2]), ip(keep[-1])
        o.event_duration, o.ev_cap = float(opt.event_duration), cap
        ev_count = torch.zeros(B, dtype=torch.int32, device=dev)
        ev_t = torch.full((B, cap), float("nan"), dtype=torch.float64, device=dev)
        ev_i = torch.full((B, cap), -1, dtype=torch.int32, device=dev)
        ev_y = torch.zeros((B, cap, n), dtype=torch.float64, device=dev)
        n_rows = n_rows_all if dense else torch.zeros(B, dtype=torch.int32, device=dev)
        out.ev_count, out.ev_t, out.ev_i, out.ev_y = ev_count.data_ptr(), ev_t.data_ptr(), ev_i.data_ptr(), ev_y.data_ptr()
        out.n_rows = n_rows.data_ptr()
    y_init = s.Y.clone()
    if callback is not None:
        return _finish_callback_events(dae, s, tspan, opt, rp, o, out, y_out, stats_d, single, dense, vars_in,
                                       on_device)
    L.check(s.lib.szb_rodas(s.h, C.byref(o), C.byref(out), s._stream()), s.h, "szb_rodas")
    st = stats_d.cpu().numpy().astype(np.int64)
    stats = Stats(opt.scheme)
    names = ("nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve")
    for i, k in enumerate(names):
        setattr(stats, k, int(st[0, i]) if single else st[:, i].copy())
    stats.flag = int(st[0, 6]) if single else st[:, 6].copy()
    failed = (st[:, 6] != 0) & (st[:, 6] != L.SZB_FLAG_EVENT)   # a terminal event is a regular end (rodas.py:267-271)
    stats.ret = ("failed" if failed[0] else None) if single else np.where(failed, "failed", None)
    # the reference never sets succeed for Rodas (stats.py:11 stays False); keep that for B = 1
    stats.succeed = False if single else ~failed
    if dense:
        Yh = y_out.permute(1, 0, 2).contiguous()          # [B, nt, n]
        T = tspan.copy()
        rows_all = n_rows_all.cpu().numpy().astype(np.int64)
        if single:
            # a scenario that stops early (dt underflow, > 100 rejects, singular matrix) returns what it reached,
            # like the reference's T[0:nt+1], Y[0:nt+1] (rodas.py:329-344) -- not zero rows
            Y = Yh[0].cpu().numpy()
            if failed[0]:
                T, Y = T[:rows_all[0]], Y[:rows_all[0]]
        else:
            short = torch.as_tensor(rows_all, device=Yh.device)[:, None] <= torch.arange(len(tspan), device=Yh.device)[None, :]
            if bool(short.any()):                          # rows a scenario never reached: nan, not zeros
                Yh[short] = float("nan")
            stats.n_rows = rows_all
            Y = Yh if on_device else Yh.cpu().numpy()
    elif single:
        nrec = int(st[0, 7])
        T = t_steps[:nrec, 0].cpu().numpy()
        Y = y_steps[:nrec, 0].cpu().numpy()
    else:
        T = np.stack([np.full(B, t0), t_end.cpu().numpy()], axis=1)
        Yh = torch.stack([y_init, s.Y], dim=1)             # [B, 2, n]
        Y = Yh.clone() if on_device else Yh.cpu().numpy()
    te = ye = ie = None
    if dev_events is not None:
        cnt = ev_count.cpu().numpy().astype(np.int64)
        stats.nevent = int(cnt[0]) if single else cnt.copy()
        if int(cnt.max(initial=0)) > dev_events.max_events:
            import warnings
            warnings.warn(f"ComponentEvents: {int(cnt.max())} events located but only max_events = "
                          f"{dev_events.max_events} are recorded per scenario; te / ye / ie hold the first ones",
                          RuntimeWarning, stacklevel=2)
        if single:
            k = min(int(cnt[0]), dev_events.max_events)
            if k > 0:
                te, ie = ev_t[0, :k].cpu().numpy(), ev_i[0, :k].cpu().numpy().astype(float)
                ye = ev_y[0, :k].cpu().numpy()
            if dense:
                # the reference's T ends at a terminal event (rodas.py:269-270, 291-299): grid nodes up to it, then the event
                rows = int(n_rows[0])
                T, Y = T[:rows], Y[:rows]
                if int(st[0, 6]) == L.SZB_FLAG_EVENT and (rows == 0 or T[-1] < te[-1]):
                    T = np.append(T, te[-1])
                    Y = np.vstack([Y, ye[-1][None, :]])
        else:                                             # [B, cap] padded with nan / -1; rows of Y past n_rows are zero
            te, ie = ev_t.cpu().numpy(), ev_i.cpu().numpy()
            ye = ev_y if on_device else ev_y.cpu().numpy()
            if not dense:
                stats.n_rows = None
    if vars_in is not None and not isinstance(Y, torch.Tensor):
        Y = TimeVars(vars_in.a, Y)
        if ye is not None and single:
            ye = TimeVars(vars_in.a, ye)
    return daesol(T, Y, te, ye, ie, stats=stats)
