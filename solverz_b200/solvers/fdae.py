"""Batched fixed-step FDAE solver with the reference's ``fdae_solver(fdae, tspan, y0, opt, **y_hist)``
surface (Solverz/solvers/fdesolver.py:14-135): every step is one batched Newton solve
(``szb_newton`` at time ``t0 + dt``) warm-started from the previous node, with the history nodes
``y_0, y_{-1}, ...`` bound as the per-scenario ``AliasVar`` parameters and refreshed on the device.
"""
from __future__ import annotations

import numpy as np

from ..num_api.model import CudaModel
from ..variables import TimeVars, Vars
from .option import Opt
from .solution import daesol
from .stats import Stats


def fdae_solver(fdae: CudaModel, tspan, y0, opt: Opt = None, **kwargs):
    import torch
    if not isinstance(fdae, CudaModel) or fdae.kind != "FDAE":
        raise TypeError("fdae_solver of solverz_b200 needs an FDAE model rendered with lang='cuda'")
    stats = Stats(scheme='FDE solver')
    if opt is None:
        opt = Opt(stats=True)
    dt = opt.step_size
    tspan = np.array(tspan, dtype=float)
    t_init, tend = tspan[0], tspan[-1]
    vars_in = y0 if isinstance(y0, Vars) else None
    y_arr = y0.array if isinstance(y0, Vars) else y0
    on_device = isinstance(y_arr, torch.Tensor)
    if not on_device:
        y_arr = np.asarray(y_arr, dtype=float)
    single = y_arr.ndim == 1
    B = 1 if single else int(y_arr.shape[0])
    s = fdae.session(B)
    s.set_params({k: v for k, v in fdae.p.items() if "_tag_" not in k})
    s.set_y(y_arr)
    nhist = fdae.nstep
    # alias parameters: name -> (tag index, slice of the base variable)
    alias = {}
    for pm in fdae.params_meta:
        nm = pm["name"]
        if "_tag_" in nm:
            base, tag = nm.rsplit("_tag_", 1)
            if base in fdae.a:
                alias[nm] = (int(tag), fdae.a[base])
    # history nodes on the device: hist[0] = y_0 (current node), hist[1] = y_{-1}, ...
    hist = [s.Y.clone()]
    for j in range(1, nhist):
        yj = kwargs[f"y{j}"]
        yj = yj.array if isinstance(yj, Vars) else yj
        tj = yj if isinstance(yj, torch.Tensor) else torch.as_tensor(np.asarray(yj, dtype=float))
        tj = tj.to(s.Y.device)
        hist.append((tj.reshape(1, -1).expand(B, -1) if tj.ndim == 1 else tj).clone())
    T = [float(t_init)]
    Ys = [s.Y.clone()]
    # backend knob for batches whose whole trajectory does not fit the device: opt.trajectory = 'last' keeps the
    # first and the last node (Y is [B, 2, n], T = [t0, tend]); the default keeps every node like the reference
    keep_all = getattr(opt, "trajectory", "all") != "last"
    tt = t0 = float(t_init)
    uround = np.spacing(1.0)
    it_sum = torch.zeros(B, dtype=torch.int64, device=s.Y.device)
    ok = torch.ones(B, dtype=torch.bool, device=s.Y.device)
    done = False
    while not done:
        if tt + dt >= tend:                                   # fdesolver.py:96-99
            dt = tend - tt
        else:
            dt = np.minimum(dt, 0.5 * (tend - tt))
        for nm, (tag, sl) in alias.items():                   # F(t0 + dt, y, p, Y[nt], Y[nt-1], ...)
            s.P[nm].copy_(hist[tag][:, sl])
        s.set_time(t0 + dt)
        # nr_method(ae, y0, Opt(ite_tol=...)), :108 -- the reference always runs it with its default max_it = 100;
        # an Opt(max_it=...) below that is honoured here (bounds the time a stagnating scenario can hold the batch)
        iters, flags = s.newton(opt.ite_tol, min(100, int(getattr(opt, "max_it", 100))))
        it_sum += iters.to(torch.int64)
        ok &= flags == 0
        tt = tt + dt
        T.append(float(tt))
        if keep_all:
            Ys.append(s.Y.clone())
        else:                                                 # opt.trajectory = 'last': first and last node only
            Ys[1:] = [s.Y.clone()]
        t0 = tt
        hist = [s.Y.clone()] + hist[:nhist - 1]
        if np.abs(tend - tt) < uround:
            done = True
    s.set_time(0.0)
    nt = len(T) - 1
    it = it_sum.cpu().numpy()
    stats.nstep = nt
    if single:
        stats.nfeval = int(it[0]) + nt
        stats.nJeval = stats.ndecomp = stats.nsolve = int(it[0])
        stats.succeed = bool(ok[0])
    else:
        stats.nfeval = it + nt
        stats.nJeval, stats.ndecomp, stats.nsolve = it.copy(), it.copy(), it.copy()
        stats.succeed = ok.cpu().numpy()
    Y = torch.stack(Ys, dim=1)                                # [B, nt+1, n]
    T = np.array(T) if keep_all else np.array([T[0], T[-1]])
    if single:
        Y = Y[0].cpu().numpy()
        if vars_in is not None:
            Y = TimeVars(vars_in.a, Y)
    elif not on_device:
        Y = Y.cpu().numpy()
    return daesol(T, Y, stats=stats)
