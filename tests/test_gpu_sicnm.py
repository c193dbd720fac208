"""SURVEY.md §8f item 3 on the GPU: ``HVP`` of a module rendered with ``make_hvp=True`` and the batched ``sicnm``
against the reference's own runs (tests/golden/make_golden.py::golden_sicnm; Solverz/solvers/nlaesolver/
sicnm.py:13-369, equation/hvp.py:12-74)."""
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"
sys.path.insert(0, str(GOLD))


def _model(tmp_path_factory, make_hvp=True):
    import solverz_b200 as szb
    from make_golden import sicnm_model
    ae, y0 = sicnm_model(szb)
    return szb.made_numerical(ae, y0, sparse=True, make_hvp=make_hvp), y0


def test_hvp_matrix(libszb, tmp_path_factory):
    g = np.load(GOLD / "sicnm_ref.npz")
    mdl, y0 = _model(tmp_path_factory)
    H = mdl.HVP(g["yp"], mdl.p, g["v"]).toarray()
    np.testing.assert_allclose(H, g["H"], rtol=0, atol=1e-13)
    np.testing.assert_allclose(mdl.J(g["yp"], mdl.p).toarray(), g["J"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("scheme", ["rodas4", "rodas3d"])
def test_sicnm_single_and_batch(libszb, tmp_path_factory, scheme):
    """One scenario at a time like the reference, then the four parameter sets as one batch: identical step /
    stage / factorisation / reject counts, solutions at 1e-10."""
    from solverz_b200 import Opt, nr_method, sicnm
    g = np.load(GOLD / "sicnm_ref.npz")
    mdl, y0 = _model(tmp_path_factory)
    A, Bp = g["A"], g["B"]
    mdl.p["a"], mdl.p["b"] = A[1], Bp[1]
    sol = sicnm(mdl, y0, Opt(ite_tol=1e-8, scheme=scheme))
    st = sol.stats
    assert [st.nstep, st.nfeval, st.ndecomp, st.nreject] == g[f"{scheme}_stats"][1].tolist()
    np.testing.assert_allclose(sol.y.array, g[f"{scheme}_y"][1], rtol=0, atol=1e-10)
    # the pseudo-time inherits the last-bit differences of the two factorisations (2n x 2n with partial pivoting
    # there, block-eliminated n x n here) through the controller's err**(1/pord); the iterates do not care
    assert abs(st.T[-1] - g[f"{scheme}_tend"][1]) <= 1e-6 * g[f"{scheme}_tend"][1] and st.succeed
    assert np.abs(mdl.F(sol.y.array, mdl.p)).max() < 1e-8
    mdl.p["a"], mdl.p["b"] = A, Bp
    solb = sicnm(mdl, np.tile(y0.array, (len(A), 1)), Opt(ite_tol=1e-8, scheme=scheme))
    sb = solb.stats
    got = np.stack([sb.nstep, sb.nfeval, sb.ndecomp, sb.nreject], axis=1)
    assert got.tolist() == g[f"{scheme}_stats"].tolist()
    np.testing.assert_allclose(solb.y, g[f"{scheme}_y"], rtol=0, atol=1e-10)
    np.testing.assert_allclose(sb.T, g[f"{scheme}_tend"], rtol=1e-6)
    assert sb.succeed.all()
    # the same roots as Newton where Newton converges
    nr = nr_method(mdl, np.tile(y0.array, (len(A), 1)), Opt(ite_tol=1e-10))
    ok = nr.stats.succeed
    assert ok.any()
    np.testing.assert_allclose(solb.y[ok], nr.y[ok], rtol=0, atol=1e-7)


def test_sicnm_needs_hvp(libszb, tmp_path_factory):
    from solverz_b200 import sicnm
    mdl, y0 = _model(tmp_path_factory, make_hvp=False)
    with pytest.raises(AttributeError):
        sicnm(mdl, y0)


def test_sicnm_polar_power_flow(libszb, tmp_path_factory):
    """LoopEqn model (the reference's Hvp rejects its matrix blocks, hvp.py:43-46): H against a central difference
    of J v, and sicnm reaching Newton's power-flow solutions for a batch of injections."""
    import scipy.sparse as sps
    from solverz_b200 import Opt, made_numerical, nr_method, sicnm
    from solverz_b200.synthetic import pf_scenarios, polar_pf_model
    eqs, y0, info = polar_pf_model(6, 6)
    mdl = made_numerical(eqs, y0, sparse=True, make_hvp=True)
    rng = np.random.default_rng(5)
    y = y0.array + 0.05 * rng.standard_normal(mdl.n)
    v = rng.standard_normal(mdl.n)
    H = mdl.HVP(y, mdl.p, v).toarray()
    e = 1e-6
    Hfd = np.zeros_like(H)
    for k in range(mdl.n):
        yp, ym = y.copy(), y.copy()
        yp[k] += e
        ym[k] -= e
        Hfd[:, k] = (mdl.J(yp, mdl.p) @ v - mdl.J(ym, mdl.p) @ v) / (2 * e)
    assert np.abs(H - Hfd).max() < 1e-6 * max(1.0, np.abs(H).max())
    B = 8
    P, Q = pf_scenarios(info, B)
    mdl.p["Pinj"], mdl.p["Qinj"] = P, Q
    Y0 = np.tile(y0.array, (B, 1))
    nr = nr_method(mdl, Y0, Opt(ite_tol=1e-10))
    sc = sicnm(mdl, Y0, Opt(ite_tol=1e-8))
    assert nr.stats.succeed.all() and sc.stats.succeed.all()
    np.testing.assert_allclose(sc.y, nr.y, rtol=0, atol=1e-6)
    assert (sc.stats.ndecomp == sc.stats.nstep + sc.stats.nreject).all()
