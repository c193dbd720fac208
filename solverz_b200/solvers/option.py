"""Solver options; field names, defaults and meaning follow Solverz/solvers/option.py:4-51.  Four backend knobs come
after the reference's fields (keyword-only, so reference call sites are unaffected):

* ``dfdt``            'module' (default): dF/dt == 0 and F0 aliased to F1 in ntrp2, what the reference's RENDERED module
                      computes (two F calls return one shared buffer); 'fd': the literal formulas,
                      what the inline made_numerical path computes;
* ``pivot_fallback``  nr_method re-runs scenarios that did not converge on a pivot order analysed from their own
                      Jacobian (the reference pivots per call; default True);
* ``trajectory``      fdae_solver: 'all' (default, every node like the reference) or 'last' (first and last node only,
                      for batches whose trajectory does not fit the device);
* ``max_steps``       Rodas: safety bound on accepted + rejected steps per scenario (0: 10^6)."""
from typing import Callable


class Opt:
    def __init__(self, atol=1e-6, rtol=1e-3, f_savety=0.9, facmax=6, fac1=0.2, fac2=6, scheme='rodas4',
                 ite_tol=1e-5, fix_h: bool = False, hinit=None, hmax=None, pbar=False, stats=False,
                 step_size=1e-3, event: Callable = None, event_duration=1e-8, partial_decompose=False,
                 ode15smaxit=4, ndfmaxit=4, normcontrol=False, numJac=False, max_it=100, profile=False, *,
                 dfdt: str = "module", pivot_fallback: bool = True, trajectory: str = "all", max_steps: int = 0):
        self.atol = atol
        self.rtol = rtol
        self.f_savety = f_savety
        self.facmax = facmax
        self.fac1 = fac1
        self.fac2 = fac2
        self.fix_h = fix_h
        self.hinit = hinit
        self.hmax = hmax
        self.scheme = scheme
        self.pbar = pbar
        self.ite_tol = ite_tol
        self.stats = stats
        self.step_size = step_size
        self.event = event
        self.event_duration = event_duration
        self.partial_decompose = partial_decompose
        self.ode15smaxit = ode15smaxit
        self.ndfmaxit = ndfmaxit
        self.normcontrol = normcontrol
        self.numJac = numJac
        self.max_it = max_it
        self.profile = profile
        if dfdt not in ("module", "fd"):
            raise ValueError("Opt(dfdt=...): 'module' or 'fd'")
        if trajectory not in ("all", "last"):
            raise ValueError("Opt(trajectory=...): 'all' or 'last'")
        self.dfdt = dfdt
        self.pivot_fallback = pivot_fallback
        self.trajectory = trajectory
        self.max_steps = max_steps
