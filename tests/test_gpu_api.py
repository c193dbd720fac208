"""GPU tests of the drop-in surface with the reference's own known answers (made_numerical, nr_method,
fdae_solver, single-scenario F/J callables, batched tensors)."""
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def test_nr_x2_minus_1(libszb):
    """reference tests/test_0.py:1-17"""
    from solverz_b200 import Eqn, Model, Opt, Var, made_numerical, nr_method
    k = np.load(GOLD / "known_answers.npz")
    m = Model()
    m.x = Var(name='x', value=2)
    m.e = Eqn(name='e', eqn=m.x ** 2 - 1)
    g, y0 = m.create_instance()
    nae = made_numerical(g, y0, sparse=True)
    sol = nr_method(nae, y0, Opt(ite_tol=1e-8))
    assert abs(sol.y['x'] - 1) <= 1e-8
    assert sol.stats.nstep == int(k["x2m1_nstep"]) and sol.stats.succeed
    assert sol.stats.nfeval == sol.stats.nstep + 1


@pytest.mark.parametrize("sign", [1, -1])
def test_mat_mul_linear(libszb, sign):
    """reference tests/test_matrix_equation.py:57-83: b - A@x = 0 -> [-0.4, 0.3]"""
    from solverz_b200 import Eqn, Mat_Mul, Model, Param, Var, made_numerical, nr_method
    k = np.load(GOLD / "known_answers.npz")
    m = Model()
    m.x = Var('x', [0, 0])
    m.b = Param('b', [0.5, 1])
    m.A = Param('A', [[1, 3], [-1, 2]], dim=2, sparse=True)
    m.eqnf = Eqn('eqnf', sign * (m.b - Mat_Mul(m.A, m.x)))
    smdl, y0 = m.create_instance()
    mdl = made_numerical(smdl, y0, sparse=True)
    sol = nr_method(mdl, y0)
    np.testing.assert_allclose(sol.y.array, np.array([-0.4, 0.3]))
    assert sol.stats.nstep == int(k["matmul_linear_nstep"])


def test_mutable_matrix_jacobian(libszb):
    """x * (A@x) - b: J = diag(A@x) + diag(x)@A (reference tests/test_matrix_equation.py:103-158)."""
    from solverz_b200 import Eqn, Mat_Mul, Model, Opt, Param, Var, made_numerical, nr_method
    k = np.load(GOLD / "known_answers.npz")
    m = Model()
    m.x = Var('x', [0.5, 0.5])
    m.A = Param('A', [[2, 1], [1, 3]], dim=2, sparse=True)
    m.b = Param('b', [4.0, 5.0])
    m.eqn = Eqn('f', m.x * Mat_Mul(m.A, m.x) - m.b)
    smdl, y0 = m.create_instance()
    mdl = made_numerical(smdl, y0, sparse=True)
    J = mdl.J(y0.array, mdl.p)
    np.testing.assert_allclose(J.toarray(), k["mutable_J0"], atol=1e-12)
    sol = nr_method(mdl, y0, Opt(ite_tol=1e-10))
    np.testing.assert_allclose(sol.y.array, k["mutable_y"], atol=1e-10)
    assert sol.stats.nstep == int(k["mutable_nstep"])


def test_multi_step_fdae(libszb):
    """reference Solverz/solvers/test/test_multi_step_fdae.py:6-24"""
    from solverz_b200 import AliasVar, Eqn, Model, Opt, Var, fdae_solver, made_numerical
    m = Model()
    m.x = Var('x', [1, 2])
    m.x0 = AliasVar('x', value=m.x.value, step=1)
    m.x1 = AliasVar('x', value=m.x.value, step=2)
    m.f1 = Eqn('f1', m.x[0] + m.x0[0] + m.x1[1])
    m.f2 = Eqn('f2', m.x[1] + m.x0[1] - m.x1[0])
    smdl, y0 = m.create_instance()
    mdl = made_numerical(smdl, y0, sparse=True)
    assert mdl.nstep == 2
    sol = fdae_solver(mdl, [0, 120], y0, Opt(step_size=60), y1=y0)
    np.testing.assert_allclose(sol.Y[1], np.array([-3., -1.]))
    np.testing.assert_allclose(sol.T, [0, 60, 120])
    # x_{k+1} = -x_k - P x_{k-1}: second step by hand
    x0, x1 = np.array([1., 2.]), np.array([-3., -1.])
    x2 = np.array([-x1[0] - x0[1], -x1[1] + x0[0]])
    np.testing.assert_allclose(sol.Y[2], x2)
    # batched: different starting nodes per scenario
    B = 5
    Y0 = np.arange(2 * B, dtype=float).reshape(B, 2) + 1
    solb = fdae_solver(mdl, [0, 120], Y0, Opt(step_size=60), y1=Y0)
    for b in range(B):
        a = Y0[b]
        s1 = np.array([-a[0] - a[1], -a[1] + a[0]])
        np.testing.assert_allclose(solb.Y[b, 1], s1, atol=1e-12)
        np.testing.assert_allclose(solb.Y[b, 2], [-s1[0] - a[1], -s1[1] + a[0]], atol=1e-12)


def test_batched_device_tensor_and_failure_flags(libszb):
    """A CUDA tensor in gives a CUDA tensor out; a scenario with no root reports max_it, others converge."""
    import torch
    from solverz_b200 import Eqn, Model, Opt, Param, Var, made_numerical, nr_method
    m = Model()
    m.x = Var('x', 2.0)
    m.c = Param('c', [1.0], per_scenario=True)
    m.e = Eqn('e', m.x ** 2 - m.c)
    g, y0 = m.create_instance()
    mdl = made_numerical(g, y0, sparse=True)
    mdl.p["c"] = np.array([[1.0], [4.0], [-1.0], [9.0]])       # x**2 = -1 has no real root
    Y = torch.full((4, 1), 2.0, dtype=torch.float64, device="cuda")
    sol = nr_method(mdl, Y, Opt(ite_tol=1e-10, max_it=20))
    assert isinstance(sol.y, torch.Tensor) and sol.y.is_cuda
    y = sol.y.cpu().numpy().ravel()
    np.testing.assert_allclose(y[[0, 1, 3]], [1.0, 2.0, 3.0], atol=1e-9)
    assert sol.stats.succeed.tolist() == [True, True, False, True]
    assert sol.stats.flag[2] in (1, 5)              # max-it (nr.py:50-52) or NaN
    assert sol.stats.nstep[2] == 21 or sol.stats.flag[2] == 5   # max_it + 1 iterations like the reference
    assert sol.stats.nstep[1] == 0                  # already at the root: zero iterations


def test_time_series_dae_vs_oracle(libszb):
    """Per-scenario TimeSeriesParam (a step disturbance) in a DAE, against the oracle Rodas."""
    from oracle.rodas import rodas as oracle_rodas
    from solverz_b200 import Eqn, Model, Ode, Opt, Param, Rodas, TimeSeriesParam, Var, made_numerical
    from solverz_b200.solvers.rodas_param import Rodas_param
    B = 6
    m = Model()
    m.w = Var('w', [0.0, 0.0])
    m.v = Var('v', [1.0, 1.0])
    m.d = Param('d', [2.0, 3.0])
    m.load = TimeSeriesParam('load', v_series=[1.0, 1.0, 1.5, 1.5], time_series=[0.0, 0.5, 0.6, 5.0],
                             per_scenario=True)
    m.f = Ode('f', -m.d * m.w + m.v - m.load, diff_var=m.w)
    m.g = Eqn('g', m.v + 0.2 * m.v ** 3 - 1.2 + m.w)
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    rng = np.random.default_rng(0)
    tsw = 0.3 + 0.4 * rng.random(B)
    TS = np.stack([np.array([0.0, t1, t1 + 0.1, 5.0]) for t1 in tsw])
    VS = np.stack([np.array([1.0, 1.0, 1.0 + 0.5 * (b + 1) / B, 1.0 + 0.5 * (b + 1) / B]) for b in range(B)])
    nd.p["load__ts_t"], nd.p["load__ts_v"] = TS, VS
    tsp = np.linspace(0, 2, 11)
    sol = Rodas(nd, tsp, np.tile(y0.array, (B, 1)), Opt(rtol=1e-5, atol=1e-8, hinit=1e-3))
    rp = Rodas_param("rodas4")
    d = np.array([2.0, 3.0])
    for b in range(B):
        def load(t):
            return np.interp(t, TS[b], VS[b]) if t < TS[b, -1] else VS[b, -1]

        def F(t, y):
            w, v = y[:2], y[2:]
            return np.concatenate([-d * w + v - load(t), v + 0.2 * v ** 3 - 1.2 + w])

        def J(t, y):
            v = y[2:]
            return sps.bmat([[sps.diags(-d), sps.identity(2)], [sps.identity(2), sps.diags(1 + 0.6 * v ** 2)]]).tocsc()
        ref = oracle_rodas(F, J, nd.M, tsp, y0.array, rp, rtol=1e-5, atol=1e-8, hinit=1e-3)
        assert int(sol.stats.nstep[b]) == ref["stats"]["nstep"]
        assert int(sol.stats.nreject[b]) == ref["stats"]["nreject"]
        assert np.max(np.abs(sol.Y[b] - ref["Y"])) < 1e-8


def test_linear_solver_seam(libszb):
    """reference Solverz/solvers/laesolver.py:11-45: solve / lu_decomposition on user matrices, one matrix and
    a batch with one pattern, against SciPy's SuperLU (the reference's own solver)."""
    import scipy.sparse.linalg as sla
    from solverz_b200 import BatchedLU, lu_decomposition, solve
    rng = np.random.default_rng(3)
    n = 60
    base = sps.random(n, n, density=0.08, random_state=5, format="csc") + sps.identity(n, format="csc") * 4.0
    base = sps.csc_matrix(base)
    base.sort_indices()
    b1 = rng.standard_normal(n)
    x = solve(base, b1)
    np.testing.assert_allclose(x, sla.spsolve(base, b1), rtol=1e-10, atol=1e-12)
    B = 7
    mats = []
    for k in range(B):
        M = base.copy()
        M.data = base.data * (1.0 + 0.2 * rng.standard_normal(base.nnz))
        M.setdiag(M.diagonal() + 3.0)
        mats.append(sps.csc_matrix(M))
    rhs = rng.standard_normal((B, n))
    lu = lu_decomposition(mats)
    assert sorted(lu.perm_r.tolist()) == list(range(n)) and sorted(lu.perm_c.tolist()) == list(range(n))
    assert lu.nnz >= base.nnz
    X = lu.solve(rhs)
    for k in range(B):
        np.testing.assert_allclose(X[k], sla.spsolve(mats[k], rhs[k]), rtol=1e-10, atol=1e-12)
    # a second right-hand side reuses the factors (the .solve(b) of the reference's sp_decomposition)
    rhs2 = rng.standard_normal((B, n))
    X2 = lu.solve(rhs2)
    np.testing.assert_allclose(X2[3], sla.spsolve(mats[3], rhs2[3]), rtol=1e-10, atol=1e-12)
    # refactor with new values on the same object, tensors in / tensors out
    import torch
    vals = torch.as_tensor(np.stack([m.data for m in mats])[::-1].copy(), device="cuda")
    Xt = lu.refactor(vals).solve(torch.as_tensor(rhs, device="cuda"))
    assert Xt.is_cuda
    np.testing.assert_allclose(Xt[0].cpu().numpy(), sla.spsolve(mats[B - 1], rhs[0]), rtol=1e-10, atol=1e-12)
    lu.close()
