"""Pivot safety of the static-pivot batched LU: the reference pivots per call (laesolver.py:11-15
-> SuperLU, partial pivoting); here one order is fixed from a representative matrix, so every result is either checked
(residual of the solve), recomputed with the scenario's own pivots, or flagged -- never silently wrong."""
import warnings

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as sla

pytestmark = pytest.mark.gpu


def _family(n=40, seed=0):
    """Matrices with one pattern: a strong diagonal in the representative one, a tiny (1e-13) or zero leading
    diagonal entry in some members -- fine for partial pivoting, fatal for the representative's pivot order."""
    rng = np.random.default_rng(seed)
    A0 = sps.random(n, n, density=0.15, random_state=seed, format="lil")
    A0.setdiag(4.0 + rng.random(n))
    for i in range(n - 1):
        A0[i, i + 1] = 1.0 + rng.random()
        A0[i + 1, i] = 1.0 + rng.random()
    A0 = sps.csc_matrix(A0)
    A0.sort_indices()
    return A0


def test_batched_lu_solve_is_checked_and_repaired(libszb):
    from solverz_b200.solvers.laesolver import BatchedLU
    A0 = _family()
    n = A0.shape[0]
    rng = np.random.default_rng(1)
    B = 6
    mats = []
    for b in range(B):
        M = A0.copy().tolil()
        M[:, :] = M            # keep the pattern
        M = sps.csc_matrix(M)
        M.data = A0.data * (1 + 0.1 * rng.standard_normal(A0.nnz))
        mats.append(M)
    # members 2 and 4: the first two diagonal entries collapse (explicitly stored, so the pattern is unchanged)
    for b, val in ((2, 1e-13), (4, 0.0)):
        M = mats[b].tolil()
        M[0, 0] = val if val != 0.0 else 1e-300
        M[1, 1] = val if val != 0.0 else 1e-300
        mats[b] = sps.csc_matrix(M)
        mats[b].sort_indices()
        assert np.array_equal(mats[b].indices, A0.indices)
    vals = np.stack([m.data for m in mats])
    rhs = rng.standard_normal((B, n))
    lu = BatchedLU(A0)
    lu.refactor(vals)
    X = lu.solve(rhs)
    assert set(lu.unstable.tolist()) >= {2, 4}, lu.unstable      # the check caught them
    for b in range(B):
        xr = sla.spsolve(mats[b], rhs[b])
        assert np.abs(X[b] - xr).max() <= 1e-9 * max(1.0, np.abs(xr).max()), b
    # L / U of the reference object (laesolver.py:37-42): A[perm_r][:, perm_c] = L U
    Lm, Um = lu.L, lu.U
    P = mats[0][lu.perm_r][:, lu.perm_c]
    assert abs(Lm @ Um - P).max() <= 1e-12 * abs(P).max()
    assert abs(sps.triu(Lm, 1)).max() == 0 and abs(sps.tril(Um, -1)).max() == 0
    lu.close()


def test_singular_matrix_behaves_like_the_reference(libszb):
    from solverz_b200.solvers.laesolver import BatchedLU, lu_decomposition
    A0 = _family(20, 3)
    S = A0.copy()
    S.data = S.data.copy()
    S = S.tolil()
    S[5, :] = 0.0                                   # a zero row: exactly singular, pattern kept through explicit zeros
    Sd = sps.csc_matrix(A0)
    Sd.data = np.array([0.0 if r == 5 else v for r, v in zip(A0.indices, A0.data)])
    lu = BatchedLU(A0)
    lu.refactor(np.stack([A0.data, Sd.data]))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        X = lu.solve(np.ones((2, 20)))
    assert np.all(np.isfinite(X[0])) and np.all(np.isnan(X[1]))
    assert any("singular" in str(x.message) for x in w)
    lu.close()
    with pytest.raises(RuntimeError, match="singular"):       # scipy.sparse.linalg.splu raises the same
        lu_decomposition(Sd)


def test_newton_pivot_fallback(libszb):
    """A scenario whose Jacobian has a vanishing entry where the representative point has its pivot: Newton on the
    model's static order diverges or flags it; the fallback re-analyses on that scenario's own Jacobian and reaches
    the solution the reference (partial pivoting) finds."""
    from solverz_b200 import Eqn, Model, Opt, Param, Var, made_numerical, nr_method
    m = Model()
    m.x = Var('x', [1.0, 1.0, 1.0])
    m.a = Param('a', [4.0], per_scenario=True)
    m.b = Param('b', [2.0, 5.0, 6.0], per_scenario=True)
    # J = [[a, 1, 0], [1, 3, 1], [0, 1, 2 x2]]: the order chosen at a = 4 pivots on a first
    m.e0 = Eqn('e0', m.a * m.x[0] + m.x[1] - m.b[0])
    m.e1 = Eqn('e1', m.x[0] + 3 * m.x[1] + m.x[2] - m.b[1])
    m.e2 = Eqn('e2', m.x[1] + m.x[2] ** 2 - m.b[2])
    eqs, y0 = m.create_instance()
    mdl = made_numerical(eqs, y0, sparse=True)
    a = np.array([[4.0], [0.0], [3.0], [1e-18]])
    bb = np.tile(np.array([2.0, 5.0, 6.0]), (4, 1))
    mdl.p["a"], mdl.p["b"] = a, bb
    Y0 = np.tile(y0.array, (4, 1))
    opt = Opt(ite_tol=1e-10)
    sol = nr_method(mdl, Y0, opt)
    assert sol.stats.succeed.all(), sol.stats.flag
    for k in range(4):
        x = sol.y[k]
        r = np.array([a[k, 0] * x[0] + x[1] - 2.0, x[0] + 3 * x[1] + x[2] - 5.0, x[1] + x[2] ** 2 - 6.0])
        assert np.abs(r).max() <= 1e-10
    assert np.allclose(sol.y[1], [-3.0, 2.0, 2.0], atol=1e-9) and int(sol.stats.nstep[1]) == 5   # numpy's Newton with pivoting
    opt2 = Opt(ite_tol=1e-10)
    opt2.pivot_fallback = False
    sol2 = nr_method(mdl, Y0, opt2)
    assert not sol2.stats.succeed[1]                # without the fallback the zero pivot is flagged, not hidden
    assert int(sol2.stats.flag[1]) in (2, 5, 1)


def test_rodas_flags_singular_scenario(libszb):
    """M - dt*gamma*J with a zero static pivot stops that scenario with SZB_FLAG_SINGULAR (the reference's splu would
    raise inside the step, rodas.py:159-162); the other scenarios of the batch finish."""
    from solverz_b200 import Eqn, Model, Ode, Opt, Param, Rodas, Var, made_numerical
    m = Model()
    m.x = Var('x', [1.0])
    m.z = Var('z', [1.0])
    m.k = Param('k', [1.0], per_scenario=True)
    m.f = Ode('f', -m.x + m.z, diff_var=m.x)
    m.g = Eqn('g', m.k * m.z - m.k * m.x)            # k = 0: the algebraic row vanishes, the pencil is singular
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    nd.p["k"] = np.array([[1.0], [0.0], [2.0]])
    sol = Rodas(nd, [0, 1], np.tile(y0.array, (3, 1)), Opt(rtol=1e-4, atol=1e-7))
    assert sol.stats.succeed[0] and sol.stats.succeed[2] and not sol.stats.succeed[1]
    assert int(sol.stats.flag[1]) == 2
