"""Oracle: the reference's Newton loop, restated line by line.

TEST INFRASTRUCTURE ONLY.  Follows Solverz/solvers/nlaesolver/nr.py:39-66 and the linear-solve seam
Solverz/solvers/laesolver.py:11-15 (``scipy.sparse.linalg.spsolve`` = SuperLU ``gssv``, COLAMD, partial
pivoting; SciPy 1.18.1 in this image, pinned ``scipy>=1.14.0`` by the reference's pyproject.toml:21).
Parity pinned by tests/golden/pf_ref_*.npz (outputs of the reference itself, see
tests/golden/make_golden.py) and by the reference's own known answers (tests/test_oracle.py).
"""
import numpy as np
import scipy.sparse.linalg as sla


def nr_method(F, J, y0, tol=1e-5, max_it=100):
    """F(y)->ndarray, J(y)->csc.  Returns dict(y, nstep, nfeval, succeed).  nr.py:39-66."""
    y = np.array(y0, dtype=float)
    nstep = 0
    df = F(y)                                   # nr.py:45
    nfeval = 1
    while np.max(np.abs(df)) > tol:             # nr.py:49
        if nstep > max_it:                      # nr.py:51 (allows max_it + 1 iterations)
            break
        nstep += 1                              # nr.py:55
        y = y - sla.spsolve(J(y), df)           # nr.py:57, laesolver.py:13
        df = F(y)                               # nr.py:61
        nfeval += 1
    succeed = bool(np.max(np.abs(df)) < tol)    # nr.py:63-64
    return dict(y=y, nstep=nstep, nfeval=nfeval, nJeval=nstep, succeed=succeed, fnorm=float(np.max(np.abs(df))))
