"""GPU parity tests of the hot path on the synthetic polar power-flow model (configs #2/#3 at
sizes the oracle finishes in seconds).  Mirrors the reference's three-way parity tests
(tests/test_matrix_equation.py:111-158: J at every Newton step 1e-12, solutions 1e-10)."""
import importlib
import sys

import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as sla

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pf8(libszb, tmp_path_factory):
    from solverz_b200 import module_printer
    from solverz_b200.synthetic import polar_pf_model
    d = tmp_path_factory.mktemp("mods")
    eqs, y0, info = polar_pf_model(8, 8)
    pkg = module_printer(eqs, y0, "pf8_test", directory=str(d)).render()
    sys.path.insert(0, str(d))
    mod = importlib.import_module("pf8_test")
    return mod.mdl, mod.y, info


def _oracle(info):
    from oracle.pf import PFOracle
    return PFOracle(info["Ybus"], info["pv"], info["pq"])


def test_F_J_parity(pf8):
    from solverz_b200.synthetic import pf_scenarios
    mdl, y0, info = pf8
    orc = _oracle(info)
    B = 48
    P, Q = pf_scenarios(info, B)
    rng = np.random.default_rng(0)
    Y = y0.array[None, :] + 0.05 * rng.standard_normal((B, mdl.n))
    s = mdl.session(B)
    s.set_params({"Pinj": P, "Qinj": Q})
    s.set_y(Y)
    F = s.eval_F().cpu().numpy()
    JV = s.eval_J().cpu().numpy()
    for b in range(B):
        Fo = orc.F(Y[b], P[b], Q[b])
        Jo = orc.J(Y[b])
        Jg = sps.csc_matrix((JV[b], mdl.indices, mdl.indptr), shape=(mdl.n, mdl.n))
        assert np.abs(F[b] - Fo).max() <= 1e-12 * max(1.0, np.abs(Fo).max())
        assert abs(Jg - Jo).max() <= 1e-12 * abs(Jo).max()


def test_single_scenario_callables(pf8):
    """mdl.F(y, p) / mdl.J(y, p) keep the reference's call shape for B = 1."""
    mdl, y0, info = pf8
    orc = _oracle(info)
    F = mdl.F(y0.array, mdl.p)
    J = mdl.J(y0.array, mdl.p)
    assert np.abs(F - orc.F(y0.array, info["Pbase"], info["Qbase"])).max() <= 1e-13
    assert abs(sps.csc_matrix(J) - orc.J(y0.array)).max() <= 1e-12 * abs(J).max()


def test_lu_refactor_solve(pf8):
    mdl, y0, info = pf8
    orc = _oracle(info)
    B = 40
    rng = np.random.default_rng(1)
    Y = y0.array[None, :] + 0.1 * rng.standard_normal((B, mdl.n))
    s = mdl.session(B)
    s.set_y(Y)
    s.eval_J()
    rhs = rng.standard_normal((B, mdl.n))
    import torch
    s.RHS.copy_(torch.as_tensor(rhs))
    s.lu_refactor(0.0)
    X = s.lu_solve(1).cpu().numpy()
    for b in range(B):
        Jo = orc.J(Y[b])
        xr = sla.spsolve(Jo, rhs[b])
        assert np.abs(X[b] - xr).max() <= 1e-10 * np.abs(xr).max()


@pytest.mark.parametrize("nx,B", [(8, 64), (16, 32)])
def test_newton_parity(libszb, tmp_path_factory, nx, B):
    from oracle.nr import nr_method as oracle_nr
    from solverz_b200 import Opt, module_printer, nr_method
    from solverz_b200.synthetic import pf_scenarios, polar_pf_model
    d = tmp_path_factory.mktemp(f"mods{nx}")
    eqs, y0, info = polar_pf_model(nx, nx)
    module_printer(eqs, y0, f"pfn{nx}", directory=str(d)).render()
    sys.path.insert(0, str(d))
    mdl = importlib.import_module(f"pfn{nx}").mdl
    orc = _oracle(info)
    P, Q = pf_scenarios(info, B)
    mdl.p["Pinj"], mdl.p["Qinj"] = P, Q
    sol = nr_method(mdl, np.tile(y0.array, (B, 1)), Opt(ite_tol=1e-8))
    assert sol.stats.succeed.all()
    for b in range(B):
        ref = oracle_nr(lambda y: orc.F(y, P[b], Q[b]), orc.J, y0.array, 1e-8, 100)
        assert np.abs(sol.y[b] - ref["y"]).max() <= 1e-10 * np.abs(ref["y"]).max()
        assert int(sol.stats.nstep[b]) == ref["nstep"]
        assert int(sol.stats.nfeval[b]) == ref["nfeval"]
        # residual of the GPU solution judged by the oracle's F
        assert np.abs(orc.F(sol.y[b], P[b], Q[b])).max() <= 1e-8


def test_large_fronts_in_global_memory(libszb, tmp_path_factory):
    """40x40 lattice (n = 3200): the separator fronts (order ~200) exceed shared memory, so the CTA-level
    fronts run out of the per-CTA global workspace; factor + forward + backward against SuperLU, and the
    Newton iteration against the oracle."""
    import torch
    from oracle.nr import nr_method as oracle_nr
    from solverz_b200 import Opt, module_printer, nr_method
    from solverz_b200.synthetic import pf_scenarios, polar_pf_model
    d = tmp_path_factory.mktemp("mods40")
    eqs, y0, info = polar_pf_model(40, 40)
    module_printer(eqs, y0, "pfn40", directory=str(d)).render()
    sys.path.insert(0, str(d))
    mdl = importlib.import_module("pfn40").mdl
    orc = _oracle(info)
    B = 6
    s = mdl.session(B)
    assert s.sym.f_top_max > 167, "this test is meant to exercise fronts larger than shared memory"
    rng = np.random.default_rng(3)
    Y = y0.array[None, :] + 0.05 * rng.standard_normal((B, mdl.n))
    s.set_y(Y)
    s.eval_J()
    rhs = rng.standard_normal((B, mdl.n))
    s.RHS.copy_(torch.as_tensor(rhs))
    s.lu_refactor(0.0)
    X = s.lu_solve(1).cpu().numpy()
    for b in range(B):
        xr = sla.spsolve(orc.J(Y[b]), rhs[b])
        assert np.abs(X[b] - xr).max() <= 1e-10 * np.abs(xr).max()
    P, Q = pf_scenarios(info, B)
    mdl.p["Pinj"], mdl.p["Qinj"] = P, Q
    sol = nr_method(mdl, np.tile(y0.array, (B, 1)), Opt(ite_tol=1e-8))
    assert sol.stats.succeed.all()
    for b in range(B):
        ref = oracle_nr(lambda y: orc.F(y, P[b], Q[b]), orc.J, y0.array, 1e-8, 100)
        assert np.abs(sol.y[b] - ref["y"]).max() <= 1e-10 * np.abs(ref["y"]).max()
        assert int(sol.stats.nstep[b]) == ref["nstep"]


@pytest.mark.parametrize("outer_nb", [8, 24])
def test_large_fronts_with_short_outer_blocks(libszb, tmp_path_factory, monkeypatch, outer_nb):
    """Fronts beyond ~1100 rows leave room for fewer than 32 side-panel columns in shared memory, so the outer blocks
    of the global-memory elimination get shorter (24, 16, 8 pivots).  The SZB_OUTER_NB test hook of lu_plan forces
    those block lengths on the 40 x 40 lattice: factor + solve against SuperLU.  (The district-heating model at
    49 612 pipes, front 950, was the first to get a block length below 32 -- 28, not a whole number of 8-row tiles --
    and found a group barrier that only some warps executed; block lengths are whole tiles since.)"""
    import torch
    from solverz_b200 import module_printer
    from solverz_b200.synthetic import polar_pf_model
    monkeypatch.setenv("SZB_OUTER_NB", str(outer_nb))
    d = tmp_path_factory.mktemp(f"mods40nb{outer_nb}")
    eqs, y0, info = polar_pf_model(40, 40)
    name = f"pfn40nb{outer_nb}"
    module_printer(eqs, y0, name, directory=str(d)).render()
    sys.path.insert(0, str(d))
    mdl = importlib.import_module(name).mdl
    orc = _oracle(info)
    B = 4
    s = mdl.session(B)
    assert s.sym.f_top_max > 167
    rng = np.random.default_rng(11)
    Y = y0.array[None, :] + 0.05 * rng.standard_normal((B, mdl.n))
    s.set_y(Y)
    s.eval_J()
    rhs = rng.standard_normal((B, mdl.n))
    s.RHS.copy_(torch.as_tensor(rhs))
    s.lu_refactor(0.0)
    assert (s.lu_status().cpu().numpy() == 0).all()
    X = s.lu_solve(1).cpu().numpy()
    for b in range(B):
        xr = sla.spsolve(orc.J(Y[b]), rhs[b])
        assert np.abs(X[b] - xr).max() <= 1e-10 * np.abs(xr).max()


def test_full_size_config2_properties(libszb, tmp_path_factory):
    """BASELINE.json configs[1] at full size (1024 buses, n = 2048, 4096 load scenarios) through properties
    that do not need 4096 oracle solves: every scenario converges; the ORACLE's residual of every GPU
    solution is below the tolerance; a sample agrees with the oracle's Newton (solution and counts); the
    batch is order-independent bit for bit (scenario b of a permuted batch = scenario perm[b])."""
    from oracle.nr import nr_method as oracle_nr
    from solverz_b200 import Opt, module_printer, nr_method
    from solverz_b200.synthetic import pf_scenarios, polar_pf_model
    d = tmp_path_factory.mktemp("mods32")
    eqs, y0, info = polar_pf_model(32, 32)
    module_printer(eqs, y0, "pfn32", directory=str(d)).render()
    sys.path.insert(0, str(d))
    mdl = importlib.import_module("pfn32").mdl
    assert mdl.n == 2048
    orc = _oracle(info)
    B = 4096
    P, Q = pf_scenarios(info, B)
    mdl.p["Pinj"], mdl.p["Qinj"] = P, Q
    Y0 = np.tile(y0.array, (B, 1))
    sol = nr_method(mdl, Y0, Opt(ite_tol=1e-8))
    assert sol.stats.succeed.all() and (sol.stats.flag == 0).all()
    assert sol.stats.nstep.min() >= 2 and sol.stats.nstep.max() <= 6
    assert (sol.stats.nfeval == sol.stats.nstep + 1).all()
    res = np.array([np.abs(orc.F(sol.y[b], P[b], Q[b])).max() for b in range(B)])
    assert res.max() <= 1e-8
    for b in range(0, B, 128):                          # 32 scenarios against the oracle's own Newton
        ref = oracle_nr(lambda y: orc.F(y, P[b], Q[b]), orc.J, y0.array, 1e-8, 100)
        assert np.abs(sol.y[b] - ref["y"]).max() <= 1e-10 * np.abs(ref["y"]).max()
        assert int(sol.stats.nstep[b]) == ref["nstep"]
    perm = np.random.default_rng(0).permutation(B)
    mdl.p["Pinj"], mdl.p["Qinj"] = P[perm], Q[perm]
    sol2 = nr_method(mdl, Y0, Opt(ite_tol=1e-8))
    assert np.array_equal(sol2.y, sol.y[perm]) and np.array_equal(sol2.stats.nstep, sol.stats.nstep[perm])


def test_full_size_config3_properties(libszb, tmp_path_factory):
    """BASELINE.json configs[2] at its per-GPU size (10 000 buses, n = 20 000, 16 384 scenarios over 8 GPUs =
    2048 per GPU; separator fronts up to 779 x 779 run out of the per-CTA global workspace): every scenario
    converges, the ORACLE's residual of every GPU solution is below the tolerance, and 16 sampled scenarios
    agree with the oracle's own Newton (solution 1e-10 relative, identical iteration counts)."""
    from oracle.nr import nr_method as oracle_nr
    from solverz_b200 import Opt, module_printer, nr_method
    from solverz_b200.synthetic import pf_scenarios, polar_pf_model
    d = tmp_path_factory.mktemp("mods100")
    eqs, y0, info = polar_pf_model(100, 100)
    module_printer(eqs, y0, "pfn100", directory=str(d)).render()
    sys.path.insert(0, str(d))
    mdl = importlib.import_module("pfn100").mdl
    assert mdl.n == 20000
    orc = _oracle(info)
    B = 2048
    P, Q = pf_scenarios(info, B)
    mdl.p["Pinj"], mdl.p["Qinj"] = P, Q
    sol = nr_method(mdl, np.tile(y0.array, (B, 1)), Opt(ite_tol=1e-8))
    assert mdl.session(B).sym.f_top_max > 500
    assert sol.stats.succeed.all() and (sol.stats.flag == 0).all()
    assert sol.stats.nstep.min() >= 2 and sol.stats.nstep.max() <= 8
    assert (sol.stats.nfeval == sol.stats.nstep + 1).all()
    res = np.array([np.abs(orc.F(sol.y[b], P[b], Q[b])).max() for b in range(B)])
    assert res.max() <= 1e-8
    for b in range(0, B, 128):                          # 16 scenarios against the oracle's own Newton
        ref = oracle_nr(lambda y: orc.F(y, P[b], Q[b]), orc.J, y0.array, 1e-8, 100)
        assert np.abs(sol.y[b] - ref["y"]).max() <= 1e-10 * np.abs(ref["y"]).max()
        assert int(sol.stats.nstep[b]) == ref["nstep"]
    mdl.close()
