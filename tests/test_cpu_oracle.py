"""CPU tests: the oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py) and against the reference's own known answers."""
from pathlib import Path

import numpy as np
import scipy.sparse as sps

GOLD = Path(__file__).parent / "golden"


def _pf(nx):
    from oracle.pf import PFOracle
    from solverz_b200.synthetic import lattice_grid
    nb, Y = lattice_grid(nx, nx)
    npv = max(1, int(0.1 * nb))
    return PFOracle(Y, np.arange(1, 1 + npv), np.arange(1 + npv, nb)), nb


def test_pf_oracle_F_J_match_reference_module():
    g = np.load(GOLD / "pf_ref_8x8.npz")
    orc, nb = _pf(8)
    rng = np.random.default_rng(1)
    Pb = rng.uniform(-.05, .05, nb)
    Qb = rng.uniform(-.02, .02, nb)
    for y, F, Jd in ((g["y0"], g["F0"], g["J0_data"]), (g["yp"], g["Fp"], g["Jp_data"])):
        assert np.abs(orc.F(y, Pb, Qb) - F).max() <= 1e-13
        Jref = sps.csc_matrix((Jd, g["J_indices"], g["J_indptr"]), shape=(2 * nb, 2 * nb))
        assert abs(orc.J(y) - Jref).max() <= 1e-12 * abs(Jref).max()


def test_newton_oracle_matches_reference_runs():
    """Iteration counts identical, solutions to 1e-12 (reference: numba F/J + spsolve)."""
    from oracle.nr import nr_method
    g = np.load(GOLD / "pf_ref_8x8.npz")
    orc, nb = _pf(8)
    for b in range(len(g["nstep"])):
        P, Q = g["P"][b], g["Q"][b]
        r = nr_method(lambda y: orc.F(y, P, Q), orc.J, g["y0"], 1e-8, 100)
        assert r["nstep"] == int(g["nstep"][b])
        assert r["nfeval"] == int(g["nfeval"][b])
        assert np.abs(r["y"] - g["sol"][b]).max() <= 1e-12
        assert r["succeed"]


def test_newton_oracle_known_answers():
    """Reference tests: x**2 - 1 -> 1 (tests/test_0.py:16-17); b - A x -> [-0.4, 0.3]
    (tests/test_matrix_equation.py:14-83); nr.py max_it off-by-one (nr.py:50-52)."""
    from oracle.nr import nr_method
    k = np.load(GOLD / "known_answers.npz")
    r = nr_method(lambda y: y ** 2 - 1, lambda y: sps.csc_matrix(np.diag(2 * y)), np.array([2.0]), 1e-8)
    assert abs(r["y"][0] - 1) <= 1e-8 and r["nstep"] == int(k["x2m1_nstep"])
    A = np.array([[1., 3.], [-1., 2.]])
    b = np.array([0.5, 1.0])
    r = nr_method(lambda y: b - A @ y, lambda y: sps.csc_matrix(-A), np.zeros(2))
    np.testing.assert_allclose(r["y"], [-0.4, 0.3])
    assert r["nstep"] == int(k["matmul_linear_nstep"])
    # a residual that never converges runs max_it + 1 iterations and reports failure
    r = nr_method(lambda y: np.array([1.0 + 0 * y[0]]), lambda y: sps.csc_matrix([[1.0]]), np.zeros(1), 1e-8, 7)
    assert r["nstep"] == 8 and not r["succeed"]


def _dae_test_fns():
    F = lambda t, y: np.array([-y[0] ** 3 + 0.5 * y[1] ** 2, y[0] ** 2 + y[1] ** 2 - 2])
    J = lambda t, y: sps.csc_matrix(np.array([[-3 * y[0] ** 2, y[1]], [2 * y[0], 2 * y[1]]]))
    M = sps.csc_matrix(np.array([[1., 0], [0, 0]]))
    return F, J, M


def test_rodas_oracle_matches_reference_golden_sheets():
    """tests/dae_test.xlsx sheets of the reference, tolerance of its own test (tests/test_dae.py:44-48)."""
    from oracle.rodas import rodas
    from solverz_b200.solvers.rodas_param import Rodas_param
    g = np.load(GOLD / "dae_test_rodas.npz")
    F, J, M = _dae_test_fns()
    r = rodas(F, J, M, [0, 20], np.array([1., 1.]), Rodas_param("rodas4"), hinit=0.1)
    assert np.abs(r["Y"] - g["rodas_sheet"]).max() < 1e-8
    assert np.abs(r["Y"] - g["ref_Y"]).max() < 1e-11
    np.testing.assert_allclose(r["T"], g["ref_T"], rtol=1e-10)
    st = r["stats"]
    assert [st[k] for k in ("nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve")] == g["ref_stats"].tolist()
    r = rodas(F, J, M, np.linspace(0, 20, 201), np.array([1., 1.]), Rodas_param("rodas4"), hinit=0.1)
    assert np.abs(r["Y"] - g["rodas_dense_sheet"]).max() < 1e-8
    assert np.abs(r["Y"] - g["ref_dense_Y"]).max() < 1e-11


def test_rodas_oracle_other_schemes():
    from oracle.rodas import rodas
    from solverz_b200.solvers.rodas_param import Rodas_param
    g = np.load(GOLD / "dae_test_rodas.npz")
    F, J, M = _dae_test_fns()
    for sch in ("rodasp", "rodas5p", "rodas3d"):
        r = rodas(F, J, M, [0, 20], np.array([1., 1.]), Rodas_param(sch), hinit=0.1)
        assert np.abs(r["Y"] - g[f"ref_{sch}_Y"]).max() < 1e-11
        st = r["stats"]
        assert [st[k] for k in ("nstep", "nfeval", "nJeval", "ndecomp", "nreject", "nsolve")] == \
            g[f"ref_{sch}_stats"].tolist()
