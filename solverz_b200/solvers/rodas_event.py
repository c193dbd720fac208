"""Rodas with event functions, one scenario (``Opt(event=callable)``).

Reference: Solverz/solvers/daesolver/rodas/rodas.py:99-107 (event set-up), 200-271 (sign-change
detection, secant start + bisection on the dense output ``ntrp``), 273-303 (dense output clipped at a
terminal event), 329-342 (``te / ye / ie`` of the returned ``daesol``).

The event function is a *Python callback* in the reference (``value, isterminal, direction =
events(t, y)`` with ``y`` a 1-D ndarray), so the step controller that consumes it runs on the host
here as well; every numerical piece of a step -- F, J, the refactorisation of ``M - dt*gamma*J`` and
the stage solves -- is a call into the device library (``szb_eval_F / szb_eval_J / szb_lu_refactor /
szb_lu_solve``) on a one-scenario session.  Batched ensembles without events go through
``szb_rodas`` (solvers/rodas.py), where the whole loop stays on the device.
"""
from __future__ import annotations

import numpy as np

from .rodas_param import Rodas_param
from .solution import daesol
from .stats import Stats

__all__ = ["rodas_with_events", "DeviceOps", "ComponentEvents"]


class ComponentEvents:
    """Event functions of the form ``value_i = y[index_i] - threshold_i`` with the reference's
    ``direction`` / ``isterminal`` semantics (rodas.py:99-107, 200-271).  The object is itself a valid
    ``events(t, y)`` callback -- it can be handed to the reference's ``Rodas`` unchanged -- and, because the
    forms are known, ``solverz_b200.Rodas`` runs the detection, the secant + bisection search on the dense
    output and the roll-back to the event time per scenario inside the step kernel (any batch size; schemes
    with ``ntrp1``).  At most 8 events (SZB_MAX_EVENTS)."""

    def __init__(self, index, threshold=0.0, direction=0, terminal=1, max_events=16):
        self.index = np.atleast_1d(np.asarray(index, dtype=np.int32))
        ne = len(self.index)
        self.threshold = np.broadcast_to(np.asarray(threshold, dtype=np.float64), (ne,)).copy()
        self.direction = np.broadcast_to(np.asarray(direction, dtype=np.int32), (ne,)).copy()
        self.terminal = np.broadcast_to(np.asarray(terminal, dtype=np.int32), (ne,)).copy()
        self.max_events = int(max_events)       # events recorded per scenario
        if ne < 1 or ne > 8:
            raise ValueError("ComponentEvents: between 1 and 8 events")

    def __call__(self, t, y):
        y = np.asarray(y)
        return y[self.index] - self.threshold, self.terminal, self.direction


class DeviceOps:
    """F, J-refresh, factor and solve of one scenario on the device session."""

    def __init__(self, sess):
        import torch
        self.s = sess
        self.torch = torch

    def F(self, t, y):
        self.s.set_y(y)
        return self.s.eval_F(float(t))[0].cpu().numpy().copy()

    def J(self, t, y):
        self.s.set_y(y)
        self.s.eval_J(float(t))          # values stay in SLOT_JV for the next refactorisation

    def factor(self, dt_gamma):
        self.s.lu_refactor(float(dt_gamma))   # A = M - dt*gamma*J (szb.h)

    def solve(self, rhs):
        self.s.RHS.copy_(self.torch.as_tensor(rhs).reshape(1, -1))
        return self.s.lu_solve(1)[0].cpu().numpy().copy()


def _ntrp1(y0, tau, dt, K, rp):
    # rodas.py:383-386
    return y0 + tau * dt * K @ (rp.b + (tau - 1) * (rp.c + tau * (rp.d + tau * rp.e)))


def _ntrp2(y0, y1, F0, F1, tau, dt):
    # rodas.py:389-393
    return ((1 - tau) * y0 + tau * y1 +
            tau * (tau - 1) * ((1 - 2 * tau) * (y1 - y0) + (tau - 1) * dt * F0 + tau * dt * F1))


def rodas_with_events(dev, M, tspan, y0, opt, dfdt_mode: bool = False):
    """``dev``: the four device operations of one scenario (``DeviceOps`` over a one-scenario
    BatchSession whose parameters are set; the tests drive the same controller with a stub to check
    the host logic without a GPU); ``y0``: consistent initial value (``dae_ic`` ran); ``M``: scipy
    mass matrix; returns ``daesol``."""
    events = opt.event
    stats = Stats(opt.scheme)
    rp = Rodas_param(opt.scheme)
    tspan = np.asarray(tspan, dtype=float)
    tend, t0 = tspan[-1], tspan[0]
    n = y0.shape[0]
    nt = 0
    t = t0
    hmin = 16 * np.spacing(t0)
    uround = np.spacing(1.0)
    T = [t0]
    Y = [y0.copy()]
    dense_output = len(tspan) > 2
    told = t0
    inext, tnext = 1, (tspan[1] if dense_output else None)
    value, isterminal, direction = events(t, y0)                      # :102-103
    stop = 0
    tevent = None
    te, ye, ie = [], [], []
    dt = 1e-6 * (tend - t0) if opt.hinit is None else opt.hinit      # :110-117
    dt = np.maximum(dt, hmin)
    dt = np.minimum(dt, opt.hmax)
    facmax = opt.facmax
    done = False
    reject = 0

    def ntrp(y0_, ynew_, told_, dt_, K_, tau_):
        if opt.scheme != 'rodas3':
            return _ntrp1(y0_, tau_, dt_, K_, rp)
        return _ntrp2(y0_, ynew_, dev.F(told_, y0_), dev.F(told_ + dt_, ynew_), tau_, dt_)   # :363-380

    while not done:
        if np.abs(dt) < uround:                                       # :127-130
            stats.ret = 'failed'
            break
        if reject > 100:                                              # :132-135
            stats.ret = 'failed'
            break
        if t + dt >= tend:                                            # :138-141
            dt = tend - t
        else:
            dt = np.minimum(dt, 0.5 * (tend - t))
        if opt.fix_h:
            dt = opt.hinit
        K = np.zeros((n, rp.s))
        if reject == 0:                                               # :150-152
            dev.J(t, y0)
            stats.nJeval += 1
        if dfdt_mode:                                                 # :154, :347-352
            tscale = np.maximum(0.1 * np.abs(t), 1e-8)
            ddt = t + np.sqrt(np.spacing(1)) * tscale - t
            dfdt0 = dt * (dev.F(t + ddt, y0) - dev.F(t, y0)) / ddt
        else:
            dfdt0 = np.zeros(n)
        rhs = dev.F(t, y0) + rp.g[0] * dfdt0
        stats.nfeval += 1
        dev.factor(dt * rp.gamma)                                     # :159
        stats.ndecomp += 1
        K[:, 0] = dev.solve(rhs)
        for j in range(1, rp.s):                                      # :165-173
            sum_1 = K @ rp.alpha[:, j]
            sum_2 = K @ rp.gammatilde[:, j]
            y1 = y0 + dt * sum_1
            rhs = dev.F(t + dt * rp.a[j], y1) + M @ sum_2 + rp.g[j] * dfdt0
            stats.nfeval += 1
            K[:, j] = dev.solve(rhs) - sum_2
        stats.nsolve += rp.s
        sum_1 = K @ (dt * rp.b)
        ynew = y0 + sum_1
        if not opt.fix_h:                                             # :177-188
            sum_2 = K @ (dt * rp.bd)
            SK = opt.atol + opt.rtol * np.abs(ynew)
            with np.errstate(invalid='ignore', divide='ignore'):
                err = np.max(np.abs((sum_1 - sum_2) / SK))
            if np.any(np.isinf(ynew)) or np.any(np.isnan(ynew)) or np.isnan(err):
                err = 1.0e6
            err = np.maximum(err, 1.0e-6)
            fac = opt.f_savety / (err ** (1 / rp.pord))
            fac = np.minimum(facmax, np.maximum(opt.fac1, fac))
            dtnew = dt * fac
        else:
            err = 1.0
            dtnew = dt

        if err <= 1.0:
            reject = 0
            told = t
            t = t + dt
            stats.nstep += 1
            # ---- events (:200-271)
            valueold = value
            value, isterminal, direction = events(t, ynew)
            value_save = value
            ff = np.where(np.asarray(value) * np.asarray(valueold) < 0)[0]
            for i in ff:
                v0 = valueold[i]
                v1 = value[i]
                detect = 1
                if direction[i] < 0 and v0 <= v1:
                    detect = 0
                if direction[i] > 0 and v0 >= v1:
                    detect = 0
                if not detect:
                    continue
                iterate = 1
                tL, tR = told, t
                if np.abs(v1 - v0) > uround:
                    tevent = told - v0 * dt / (v1 - v0)               # secant start, :223
                else:
                    iterate = 0
                    tevent = t
                    ynext = ynew
                tol = 128 * np.max([np.spacing(told), np.spacing(t)])
                tol = np.min([tol, np.abs(t - told)])
                while iterate > 0:
                    iterate += 1
                    tau = (tevent - told) / dt
                    ynext = ntrp(y0, ynew, told, dt, K, tau)
                    value, isterminal, direction = events(tevent, ynext)
                    if v1 * value[i] < 0:
                        tL = tevent
                        tevent = 0.5 * (tevent + tR)
                        v0 = value[i]
                    elif v0 * value[i] < 0:
                        tR = tevent
                        tevent = 0.5 * (tL + tevent)
                        v1 = value[i]
                    else:
                        iterate = 0
                    if (tR - tL) < tol:
                        iterate = 0
                    if iterate > 100:
                        print(f"Lost Event in interval [{told}, {t}].\n")
                        break
                if np.abs(tevent - told) < opt.event_duration:        # :257-259
                    break
                t = tevent
                ynew = ynext
                te.append(tevent)
                ie.append(i)
                ye.append(np.array(ynext, dtype=float))
                value, isterminal, direction = events(tevent, ynext)
                value = value_save
                if isterminal[i]:                                     # :267-271
                    if dense_output and tnext >= tevent:
                        tnext = tevent
                    stop = 1
                    break
            # ---- output (:273-309)
            if dense_output:
                while t >= tnext > told:
                    tau = (tnext - told) / dt
                    ynext = ntrp(y0, ynew, told, dt, K, tau)
                    nt += 1
                    T.append(tnext)
                    Y.append(ynext)
                    if stop and tnext >= tevent:
                        break
                    inext += 1
                    if inext <= len(tspan) - 1:
                        tnext = tspan[inext]
                        if stop and tnext >= tevent:
                            tnext = tevent
                    else:
                        tnext = tend + dt
            else:
                nt += 1
                T.append(t)
                Y.append(ynew.copy())
            if np.abs(tend - t) < uround or stop:                     # :318-319
                done = True
            y0 = ynew
            facmax = opt.fac2
        else:
            reject += 1
            stats.nreject += 1
            facmax = 1
        dt = np.min([opt.hmax, np.max([hmin, dtnew])])                # :327

    T = np.array(T)
    Y = np.array(Y)
    if te:
        return daesol(T, Y, np.array(te), np.array(ye), np.array(ie, dtype=float), stats)
    return daesol(T, Y, stats=stats)
