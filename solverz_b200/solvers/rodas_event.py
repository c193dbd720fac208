"""Event functions for the batched Rodas (``Opt(event=...)``).

Reference behaviour: Solverz/solvers/daesolver/rodas/rodas.py:99-107 (``value, isterminal, direction =
events(t, y)``), 200-271 (what an event does to a step), 273-303 (dense output up to a terminal event), 329-342
(``te / ye / ie`` of the returned ``daesol``).

Two kinds of event function:

* ``ComponentEvents`` -- thresholds on components of ``y``.  Their form is known, so detection, root search and the
  roll-back to the event time run per scenario inside the step kernel (``csrc/szb_rodas.cu``).
* any Python callable -- the reference's own interface.  A callback can only be asked on the host, so the device
  integrates ONE accepted step per running scenario (``szb_rodas`` with ``stop_after_accepted = 1``, resumed call
  after call: the whole step -- J, F, the refactorisation of ``M - dt*gamma*J``, the stage solves, the error
  controller -- stays in ``szb_rodas``), then ``EventLocator`` below asks the callback about the new states of all
  scenarios, brackets sign changes, finds the event time on the step's dense output (the stage vectors are read
  from the device), and moves a scenario back to its event (``szb_rodas_state_set`` + its row of ``SLOT_Y``) before
  the next call.  Any batch size; the cost is one host round trip per accepted step, which the callback needs anyway.
"""
from __future__ import annotations

import numpy as np

from .rodas_param import Rodas_param
from .solution import daesol
from .stats import Stats

__all__ = ["ComponentEvents", "EventLocator", "rodas_callback_events"]


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


def _ntrp1(y0, tau, dt, K, rp):
    """Dense output of the schemes with interpolation coefficients (rodas.py:383-386); K: [n, s]."""
    return y0 + tau * dt * K @ (rp.b + (tau - 1) * (rp.c + tau * (rp.d + tau * rp.e)))


def _ntrp2(y0, y1, F0, F1, tau, dt):
    """Cubic Hermite output of rodas3 (rodas.py:389-393)."""
    return ((1 - tau) * y0 + tau * y1 +
            tau * (tau - 1) * ((1 - 2 * tau) * (y1 - y0) + (tau - 1) * dt * F0 + tau * dt * F1))


class EventLocator:
    """What the reference does with ``events`` after one accepted step of one scenario (rodas.py:200-271), as a
    function of that step alone: ``after_step`` gets the step ``[told, told + dt]`` with its end states and a
    callable ``interp(tau)`` for the dense output, and returns where the scenario continues from.

    Semantics kept exactly (they decide which later sign changes are seen): only components whose value changed
    sign over the step and moves in the watched ``direction`` are searched, in index order; the first guess is the
    secant point, then the bracket is bisected until it is shorter than 128 ulp of the step ends; an event closer
    than ``event_duration`` to the step start ends the search of this step; after an event the step ENDS at the
    event (later components are tested against the shortened step), and the values remembered for the next step are
    those at the ORIGINAL step end."""

    def __init__(self, events, event_duration: float, t0: float, y0: np.ndarray):
        self.events = events
        self.event_duration = float(event_duration)
        self.value = np.asarray(events(t0, y0)[0], dtype=float)
        self.te, self.ye, self.ie = [], [], []
        self.terminal = False

    def after_step(self, told: float, dt: float, t_end: float, y_end: np.ndarray, interp):
        """Returns ``(t, y, moved)``: the time / state the scenario continues from and whether they differ from the
        step end."""
        before = self.value
        now, isterminal, direction = self.events(t_end, y_end)
        now = np.asarray(now, dtype=float)
        at_step_end = now
        t, y, moved = t_end, y_end, False
        eps = np.spacing(1.0)
        for i in np.flatnonzero(now * before < 0):
            lo_v, hi_v = before[i], now[i]
            if (direction[i] < 0 and lo_v <= hi_v) or (direction[i] > 0 and lo_v >= hi_v):
                continue
            lo_t, hi_t = told, t
            width_tol = min(128 * max(np.spacing(told), np.spacing(t)), abs(t - told))
            if abs(hi_v - lo_v) > eps:
                guess = told - lo_v * dt / (hi_v - lo_v)
                y_guess = y
                for _ in range(100):
                    y_guess = interp((guess - told) / dt)
                    now, isterminal, direction = self.events(guess, y_guess)
                    mid_v = now[i]
                    if hi_v * mid_v < 0:
                        lo_t, lo_v, guess = guess, mid_v, 0.5 * (guess + hi_t)
                    elif lo_v * mid_v < 0:
                        hi_t, hi_v, guess = guess, mid_v, 0.5 * (lo_t + guess)
                    else:
                        break
                    if hi_t - lo_t < width_tol:
                        break
                else:
                    print(f"Lost Event in interval [{told}, {t}].\n")
            else:
                guess, y_guess = t, y
            if abs(guess - told) < self.event_duration:
                at_step_end = now          # the reference leaves `value` at its last evaluation here (rodas.py:257-259)
                break
            t, y, moved = guess, y_guess, True
            self.te.append(guess)
            self.ie.append(int(i))
            self.ye.append(np.array(y_guess, dtype=float))
            isterminal = self.events(guess, y_guess)[1]
            now = at_step_end
            if isterminal[i]:
                self.terminal = True
                break
        self.value = at_step_end
        return t, y, moved


def rodas_callback_events(dae, sess, tspan, opt, rp, run_call, single: bool, vars_in):
    """Drives ``szb_rodas`` one accepted step at a time for ``Opt(event=callable)`` (module docstring).

    ``run_call(resume)`` performs one ``szb_rodas`` call with ``stop_after_accepted = 1`` and returns the tensors it
    writes (``solvers/rodas.py`` owns the option / output structs); this function owns the event logic and the
    assembly of ``T / Y / te / ye / ie`` in the reference's layout."""
    import ctypes as C
    import torch
    from ..num_api import lib as L
    B, n = sess.B, dae.n
    dense = len(tspan) > 2
    t0 = float(tspan[0])
    Y0 = sess.Y.cpu().numpy()
    loc = [EventLocator(opt.event, opt.event_duration, t0, Y0[k].copy()) for k in range(B)]
    T_steps = [[t0] for _ in range(B)]
    Y_steps = [[Y0[k].copy()] for k in range(B)]
    y_prev = Y0.copy()
    nstep_prev = np.zeros(B, dtype=np.int64)
    tt = np.zeros(B)
    told = np.zeros(B)
    dts = np.zeros(B)
    inext = np.zeros(B, dtype=np.int32)
    stats = np.zeros((B, 8), dtype=np.int32)
    out = None
    resume = False
    probe = [None]
    while True:
        out = run_call(resume)
        resume = True
        L.check(sess.lib.szb_rodas_state_get(sess.h, tt.ctypes.data, told.ctypes.data, dts.ctypes.data,
                                             inext.ctypes.data, stats.ctypes.data), sess.h, "szb_rodas_state_get")
        stepped = np.flatnonzero(stats[:, 0] > nstep_prev)
        if len(stepped):
            Yh = sess.Y.cpu().numpy()
            K_host = None
            F_ends = None
            for b in stepped.tolist():
                y_start, y_end = y_prev[b], Yh[b].copy()

                def interp(tau, b=b, y_start=y_start, y_end=y_end):
                    nonlocal K_host, F_ends
                    if opt.scheme != 'rodas3':
                        if K_host is None:            # stage vectors of the step just taken: [s, B, n] on the device
                            buf = torch.empty((rp.s, B, n), dtype=torch.float64, device=sess.Y.device)
                            L.check(sess.lib.szb_rodas_stages(sess.h, C.c_void_p(buf.data_ptr()), sess._stream()),
                                    sess.h, "szb_rodas_stages")
                            K_host = buf.cpu().numpy()
                        return _ntrp1(y_start, tau, dts[b], K_host[:, b, :].T, rp)
                    if F_ends is None:
                        F_ends = {}
                    if b not in F_ends:               # rodas3: Hermite data = F at both ends of the step
                        if probe[0] is None:          # a private one-scenario session (model.session() would retire `sess`)
                            from ..num_api.model import BatchSession
                            probe[0] = BatchSession(dae, 1, sess.device, with_lu=False)
                        one = probe[0]
                        one.set_params({k: (v[b] if (v.ndim == 2 and v.shape[0] == B) else v) for k, v in sess.P.items()})
                        one.set_y(y_start)
                        f0 = one.eval_F(float(told[b]))[0].cpu().numpy().copy()
                        one.set_y(y_end)
                        f1 = one.eval_F(float(told[b] + dts[b]))[0].cpu().numpy().copy()
                        F_ends[b] = (f0, f1)
                    return _ntrp2(y_start, y_end, F_ends[b][0], F_ends[b][1], tau, dts[b])

                t_b, y_b, moved = loc[b].after_step(float(told[b]), float(dts[b]), float(tt[b]), y_end, interp)
                if moved:
                    sess.Y[b].copy_(torch.as_tensor(y_b))
                    nxt = int(np.searchsorted(tspan, t_b, side='right')) if dense else -1
                    flag = L.SZB_FLAG_EVENT if loc[b].terminal else L.SZB_FLAG_RUNNING
                    L.check(sess.lib.szb_rodas_state_set(sess.h, b, float(t_b), nxt, flag), sess.h, "szb_rodas_state_set")
                    stats[b, 6] = flag
                    Yh[b] = y_b
                T_steps[b].append(t_b)
                Y_steps[b].append(np.array(y_b, dtype=float))
            y_prev = Yh
            nstep_prev = stats[:, 0].astype(np.int64)
        if not (stats[:, 6] == L.SZB_FLAG_RUNNING).any():
            break
    if probe[0] is not None:
        probe[0].close()
    return out, stats, loc, T_steps, Y_steps
