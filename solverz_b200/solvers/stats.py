"""Counters returned on ``sol.stats`` (Solverz/solvers/stats.py:1-23).  In batched solves every
counter is an int array of length B; ``flag`` carries the per-scenario SZB_FLAG_* status."""

_COUNTERS = ("nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve")


class Stats:
    """``scheme`` + the six work counters (zero until a solver fills them) + ``succeed`` / ``ret`` (outcome),
    ``order_variation`` (multistep schemes) and ``flag`` (batched solves)."""

    def __init__(self, scheme=None):
        self.scheme = scheme
        for k in _COUNTERS:
            setattr(self, k, 0)
        self.succeed, self.ret = False, None
        self.order_variation = self.flag = None

    def __repr__(self):
        return (f"Stats({self.scheme}: succeed={self.succeed}, "
                + ", ".join(f"{k}={getattr(self, k)}" for k in _COUNTERS) + ")")
