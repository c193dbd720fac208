"""Solution containers (Solverz/solvers/solution.py:10-99)."""


class aesol:
    def __init__(self, y, stats):
        self.y = y
        self.stats = stats


class daesol:
    def __init__(self, T=None, Y=None, te=None, ye=None, ie=None, stats=None):
        self.T = T
        self.Y = Y
        self.te = te
        self.ye = ye
        self.ie = ie
        self.stats = stats
