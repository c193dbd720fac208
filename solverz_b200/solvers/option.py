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
import inspect
from typing import Callable

_SCHEMES_DFDT = ("module", "fd")
_TRAJECTORIES = ("all", "last")


class Opt:
    """Every keyword becomes an attribute of the same name (what the solvers read: ``opt.rtol``, ``opt.event`` ...)."""

    def __init__(self, atol=1e-6, rtol=1e-3, f_savety=0.9, facmax=6, fac1=0.2, fac2=6, scheme='rodas4',
                 ite_tol=1e-5, fix_h: bool = False, hinit=None, hmax=None, pbar=False, stats=False,
                 step_size=1e-3, event: Callable = None, event_duration=1e-8, partial_decompose=False,
                 ode15smaxit=4, ndfmaxit=4, normcontrol=False, numJac=False, max_it=100, profile=False, *,
                 dfdt: str = "module", pivot_fallback: bool = True, trajectory: str = "all", max_steps: int = 0):
        given = locals()
        if dfdt not in _SCHEMES_DFDT:
            raise ValueError("Opt(dfdt=...): 'module' or 'fd'")
        if trajectory not in _TRAJECTORIES:
            raise ValueError("Opt(trajectory=...): 'all' or 'last'")
        for field in _FIELDS:
            setattr(self, field, given[field])

    def __repr__(self):
        changed = {k: getattr(self, k) for k in _FIELDS if getattr(self, k) is not _DEFAULTS[k]
                   and getattr(self, k) != _DEFAULTS[k]}
        return "Opt(" + ", ".join(f"{k}={v!r}" for k, v in changed.items()) + ")"


_DEFAULTS = {k: v.default for k, v in inspect.signature(Opt.__init__).parameters.items() if k != "self"}
_FIELDS = tuple(_DEFAULTS)
