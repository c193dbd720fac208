"""Counters returned on ``sol.stats`` (Solverz/solvers/stats.py:1-23).  In batched solves every
counter is an int array of length B; ``flag`` carries the per-scenario SZB_FLAG_* status."""


class Stats:

    def __init__(self, scheme=None):
        self.scheme = scheme
        self.nstep = 0
        self.nfeval = 0
        self.nJeval = 0
        self.ndecomp = 0
        self.nreject = 0
        self.nsolve = 0
        self.ret = None
        self.succeed = False
        self.order_variation = None
        self.flag = None

    def __repr__(self):
        fields = ("succeed", "nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve")
        return f"Stats({self.scheme}: " + ", ".join(f"{k}={getattr(self, k)}" for k in fields) + ")"
