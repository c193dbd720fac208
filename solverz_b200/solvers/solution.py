"""Solution containers with the reference's attributes and operations (Solverz/solvers/solution.py:10-99):
``aesol(y, stats)``; ``daesol(T, Y, te, ye, ie, stats)`` with ``sol[i]`` / ``sol[a:b]`` (a window of the trajectory
with the events that fall inside it) and ``sol.append(other)`` (restart chains: trajectories, events and counters are
concatenated / added).  ``Y`` / ``ye`` may be numpy arrays or ``TimeVars``."""
from __future__ import annotations

import copy

import numpy as np

from .stats import Stats

_COUNTERS = ("nstep", "nfeval", "ndecomp", "nJeval", "nsolve", "nreject")


def _cat(a, b):
    """Concatenate two trajectories along time: TimeVars know how to append themselves, arrays are stacked."""
    if hasattr(a, "append") and not isinstance(a, np.ndarray):
        a.append(b)
        return a
    return np.concatenate([np.asarray(a), np.asarray(b.array if hasattr(b, "array") else b)])


class aesol:
    __slots__ = ("y", "stats")

    def __init__(self, y, stats=None):
        self.y = y
        self.stats = stats

    def __repr__(self):
        return f"ae solution using {self.stats.scheme}, succeed: {self.stats.succeed}"


class daesol:
    __slots__ = ("T", "Y", "te", "ie", "ye", "stats")

    def __init__(self, T=None, Y=None, te=None, ye=None, ie=None, stats=None):
        self.T = T
        self.Y = Y
        self.te = te
        self.ye = ye
        self.ie = ie
        self.stats = Stats() if stats is None else stats

    def __getitem__(self, item):
        if isinstance(item, (int, np.integer)):
            return daesol(self.T[int(item)], self.Y[int(item)])
        if not isinstance(item, slice):
            raise NotImplementedError(f"Index type {type(item)} not implemented!")
        T, Y = self.T[item], self.Y[item]
        te = ye = ie = None
        if self.ie is not None:
            inside = np.flatnonzero((T[0] <= self.te) & (self.te <= T[-1])) if len(T) else np.zeros(0, dtype=int)
            if inside.size:
                te, ie = self.te[inside], self.ie[inside]
                ye = self.ye[slice(int(inside[0]), int(inside[-1]) + 1)]
            else:
                te, ye, ie = self.te[0:0], self.ye[0:0], self.ie[0:0]
        stats = copy.deepcopy(self.stats)
        if getattr(stats, "order_variation", None) is not None:
            stats.order_variation = stats.order_variation[item]
        return daesol(T, Y, te, ye, ie, stats)

    def append(self, sol: "daesol"):
        if self.stats.scheme is None:
            self.stats.scheme = sol.stats.scheme
        for k in _COUNTERS:
            setattr(self.stats, k, getattr(self.stats, k) + getattr(sol.stats, k))
        if getattr(self.stats, "order_variation", None) is not None:
            self.stats.order_variation = np.concatenate((self.stats.order_variation, sol.stats.order_variation))
        self.T = sol.T if self.T is None else np.concatenate([self.T, sol.T])
        self.Y = sol.Y if self.Y is None else _cat(self.Y, sol.Y)
        if sol.te is not None:
            if self.te is None:
                self.te, self.ye, self.ie = np.copy(sol.te), copy.deepcopy(sol.ye), np.copy(sol.ie)
            else:
                self.te = np.concatenate([self.te, sol.te])
                self.ye = _cat(self.ye, sol.ye)
                self.ie = np.concatenate([self.ie, sol.ie])
