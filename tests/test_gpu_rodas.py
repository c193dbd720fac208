"""GPU parity of the batched Rodas against the oracle and the reference's own golden trajectories
(tests/dae_test.xlsx sheets ``rodas`` / ``rodas_dense`` of the reference, 1e-8: tests/test_dae.py:44-53)."""
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sps

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _dae_test_model():
    from solverz_b200 import Eqn, Model, Ode, Var, made_numerical
    m = Model()
    m.x = Var('x', 1)
    m.y = Var('y', 1)
    m.f = Ode(name='f', f=-m.x ** 3 + 0.5 * m.y ** 2, diff_var=m.x)
    m.g = Eqn(name='g', eqn=m.x ** 2 + m.y ** 2 - 2)
    dae, y0 = m.create_instance()
    return made_numerical(dae, y0, sparse=True), y0


def test_rodas_reference_golden(libszb):
    from solverz_b200 import Opt, Rodas
    g = np.load(GOLD / "dae_test_rodas.npz")
    nd, y0 = _dae_test_model()
    sol = Rodas(nd, [0, 20], y0, opt=Opt(hinit=0.1))
    assert np.max(np.abs(g["rodas_sheet"] - sol.Y.array)) < 1e-8
    np.testing.assert_allclose(sol.T, g["ref_T"], rtol=1e-10)
    got = [sol.stats.nstep, sol.stats.nfeval, sol.stats.nJeval, sol.stats.ndecomp, sol.stats.nreject,
           sol.stats.nsolve]
    assert got == g["ref_stats"].tolist()
    sol2 = Rodas(nd, np.linspace(0, 20, 201), y0, Opt(hinit=0.1))
    assert np.max(np.abs(g["rodas_dense_sheet"] - sol2.Y.array)) < 1e-8


@pytest.mark.parametrize("scheme", ["rodasp", "rodas5p", "rodas3d"])
def test_rodas_schemes_vs_reference(libszb, scheme):
    from solverz_b200 import Opt, Rodas
    g = np.load(GOLD / "dae_test_rodas.npz")
    nd, y0 = _dae_test_model()
    sol = Rodas(nd, [0, 20], y0, opt=Opt(hinit=0.1, scheme=scheme))
    assert sol.Y.array.shape == g[f"ref_{scheme}_Y"].shape
    assert np.max(np.abs(g[f"ref_{scheme}_Y"] - sol.Y.array)) < 1e-8
    got = [sol.stats.nstep, sol.stats.nfeval, sol.stats.nJeval, sol.stats.ndecomp, sol.stats.nreject,
           sol.stats.nsolve]
    assert got == g[f"ref_{scheme}_stats"].tolist()


def test_rodas_batched_vs_oracle(libszb):
    """Stiff ensemble: B copies of a small nonlinear DAE with per-scenario parameters and explicit time
    dependence; every scenario against the oracle (same accepted / rejected step counts)."""
    from oracle.rodas import rodas as oracle_rodas
    from solverz_b200 import Eqn, Model, Ode, Opt, Param, Rodas, Var, made_numerical, t_sym, sin
    from solverz_b200.solvers.rodas_param import Rodas_param
    B = 12
    rng = np.random.default_rng(3)
    kk = 50.0 + 100.0 * rng.random((B, 1))
    amp = 0.5 + rng.random((B, 1))
    m = Model()
    m.x = Var('x', [1.0, 0.5])
    m.z = Var('z', [1.0, 0.25])
    m.k = Param('k', [100.0], per_scenario=True)
    m.a = Param('a', [1.0], per_scenario=True)
    m.f = Ode('f', -m.k * (m.x - m.a * sin(t_sym)) + m.z, diff_var=m.x)
    m.g = Eqn('g', m.z + 0.1 * m.a * m.z ** 3 - m.x ** 2 - 0.1 * m.a * m.x ** 6)
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    nd.p["k"], nd.p["a"] = kk, amp
    Y0 = np.tile(y0.array, (B, 1))
    tsp = np.linspace(0, 2, 21)
    for mode in ("module", "fd"):
        opt = Opt(rtol=1e-4, atol=1e-7)
        opt.dfdt = mode
        sol = Rodas(nd, tsp, Y0, opt)
        rp = Rodas_param("rodas4")
        M = nd.M
        for b in range(B):
            k_, a_ = kk[b, 0], amp[b, 0]

            def F(t, y):
                x, z = y[:2], y[2:]
                return np.concatenate([-k_ * (x - a_ * np.sin(t)) + z,
                                       z + 0.1 * a_ * z ** 3 - x ** 2 - 0.1 * a_ * x ** 6])

            def J(t, y):
                x, z = y[:2], y[2:]
                return sps.bmat([[sps.diags(np.full(2, -k_)), sps.identity(2)],
                                 [sps.diags(-2 * x - 0.6 * a_ * x ** 5), sps.diags(1 + 0.3 * a_ * z ** 2)]]).tocsc()
            ref = oracle_rodas(F, J, M, tsp, y0.array, rp, rtol=1e-4, atol=1e-7, dfdt_mode=1 if mode == "fd" else 0)
            assert int(sol.stats.nstep[b]) == ref["stats"]["nstep"], (mode, b)
            assert int(sol.stats.nreject[b]) == ref["stats"]["nreject"], (mode, b)
            assert int(sol.stats.nfeval[b]) == ref["stats"]["nfeval"]
            assert int(sol.stats.nJeval[b]) == ref["stats"]["nJeval"]
            assert ref["stats"]["nstep"] > 5
            # 'fd' divides a difference of two F values by ddt ~ 1e-10: one-ulp differences between
            # CUDA's and numpy's sin() are amplified to ~1e-7 there; the module path has no such term
            tol = 1e-8 if mode == "module" else 5e-5   # still below rtol = 1e-4
            assert np.max(np.abs(sol.Y[b] - ref["Y"])) < tol, (mode, b)


def test_rodas_swing_ensemble_vs_reference_golden(libszb):
    """Config #4 in miniature: 9-machine swing / flux-decay / exciter DAE (n = 90, Mat_Mul network, per-scenario
    bus fault through ``fmask`` + a TimeSeriesParam) integrated as ONE batch; every scenario against the
    trajectory the reference's own rendered module + Rodas produced (tests/golden/swing_ref_3x3.npz).  The
    reference takes 1000-2000 steps per scenario here (its module path has dF/dt = 0): accepted / rejected step
    counts, F and J evaluation counts must be identical, the trajectories must agree far below rtol."""
    from pathlib import Path
    from solverz_b200 import Opt, Rodas, made_numerical
    from solverz_b200.synthetic import swing_dae_model, swing_fault_scenarios
    g = np.load(Path(__file__).parent / "golden" / "swing_ref_3x3.npz")
    m, info = swing_dae_model(3, 3)
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    B = g["Y"].shape[0]
    fmask, tst, tsv = swing_fault_scenarios(info["nb"], B)
    nd.p["fmask"] = fmask
    nd.p["gf__ts_t"], nd.p["gf__ts_v"] = tst, tsv
    opt = Opt(rtol=1e-4, atol=1e-7, hinit=1e-4)
    opt.dfdt = "module"
    sol = Rodas(nd, g["tspan"], np.tile(y0.array, (B, 1)), opt)
    for b in range(B):
        nstep, nfeval, nJ, ndec, nrej, nsolve = (int(v) for v in g["stats"][b])
        assert int(sol.stats.nstep[b]) == nstep, (b, int(sol.stats.nstep[b]), nstep)
        assert int(sol.stats.nreject[b]) == nrej, (b, int(sol.stats.nreject[b]), nrej)
        assert int(sol.stats.nfeval[b]) == nfeval and int(sol.stats.nJeval[b]) == nJ
        d = np.abs(sol.Y[b] - g["Y"][b]).max(axis=1)
        # an output time that falls on a 1 ms fault ramp (knots of the time series) sits on a slope of
        # ~3e3 / s: there the two runs may differ by rounding-level time offsets times that slope
        on_ramp = np.array([np.min(np.abs(t - tst[b][1:5])) < 2e-3 for t in g["tspan"]])
        assert d[~on_ramp].max() < 1e-6, (b, d)
        assert d.max() < 1e-4 * max(1.0, np.abs(g["Y"][b]).max()), (b, d)


def test_dae_ic_inconsistent_start_vs_oracle(libszb):
    """DaeIc (daeic.py:24-68) from an INCONSISTENT y0: swing-type toy whose differential rows carry Jacobian
    entries far larger than their mass entry (J[delta, omega] = 314), so a value-based static matching would
    pivot those rows off the mass entries and the row-masked matrix of the initialisation would be singular.
    Every scenario of the batch against the oracle's daeic (Newton on J[AlgEqn, AlgVar] with the three-probe
    halving line search), then a Rodas run from the same inconsistent start against the oracle's Rodas."""
    from oracle.rodas import daeic as oracle_daeic, rodas as oracle_rodas
    from solverz_b200 import Eqn, Model, Ode, Opt, Param, Rodas, Var, cos, made_numerical, sin
    from solverz_b200.solvers.rodas import dae_ic
    from solverz_b200.solvers.rodas_param import Rodas_param
    B = 5
    pm = np.linspace(0.5, 0.9, B).reshape(B, 1)
    m = Model()
    m.delta = Var('delta', [0.3])
    m.omega = Var('omega', [1.0])
    m.z = Var('z', [1.0, 1.0])                       # inconsistent on purpose
    m.Pm = Param('Pm', [0.8], per_scenario=True)
    m.f1 = Ode('f1', 314 * (m.omega - 1), diff_var=m.delta)
    m.f2 = Ode('f2', (m.Pm - m.z[0] * sin(m.delta) - 2 * (m.omega - 1)) / 5, diff_var=m.omega)
    m.g1 = Eqn('g1', m.z[0] - 1.2 + 0.3 * m.z[1] * cos(m.delta))
    m.g2 = Eqn('g2', m.z[1] ** 2 + m.z[0] * m.z[1] - 2)
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    nd.p["Pm"] = pm
    s = nd.session(B)
    assert sorted(zip(*[list(map(int, a)) for a in nd.mass])) == [(0, 0), (1, 1)]
    sym = s.sym
    # the differential rows are pivoted on their mass entries whatever the values say
    assert int(sym.rowp[list(sym.colp).index(0)]) == 0 and int(sym.rowp[list(sym.colp).index(1)]) == 1
    s.set_params(nd.p)
    s.set_y(np.tile(y0.array, (B, 1)))
    dae_ic(s, 0.0, 1e-3)
    got = s.Y.cpu().numpy()
    M = nd.M

    def FJ(b):
        p = pm[b, 0]

        def F(t, y):
            d, w, z0, z1 = y
            return np.array([314 * (w - 1), (p - z0 * np.sin(d) - 2 * (w - 1)) / 5,
                             z0 - 1.2 + 0.3 * z1 * np.cos(d), z1 ** 2 + z0 * z1 - 2])

        def J(t, y):
            d, w, z0, z1 = y
            return sps.csc_matrix(np.array([[0, 314, 0, 0],
                                            [-z0 * np.cos(d) / 5, -2 / 5, -np.sin(d) / 5, 0],
                                            [-0.3 * z1 * np.sin(d), 0, 1, 0.3 * np.cos(d)],
                                            [0, 0, z1, 2 * z1 + z0]]))
        return F, J
    for b in range(B):
        F, J = FJ(b)
        ref = oracle_daeic(F, J, M, y0.array, 0.0, 1e-3)
        assert np.linalg.norm(F(0.0, y0.array)[2:]) > 1e-2          # the start really is inconsistent
        assert np.abs(got[b] - ref).max() <= 1e-12, (b, got[b], ref)
    tsp = np.linspace(0, 1, 11)
    sol = Rodas(nd, tsp, np.tile(y0.array, (B, 1)), Opt(rtol=1e-4, atol=1e-7))
    rp = Rodas_param("rodas4")
    for b in range(B):
        F, J = FJ(b)
        yi = oracle_daeic(F, J, M, y0.array, 0.0, 1e-4)
        ref = oracle_rodas(F, J, M, tsp, yi, rp, rtol=1e-4, atol=1e-7, dfdt_mode=0)
        assert int(sol.stats.nstep[b]) == ref["stats"]["nstep"] and int(sol.stats.nreject[b]) == ref["stats"]["nreject"]
        assert np.max(np.abs(sol.Y[b] - ref["Y"])) < 1e-8


def test_config4_full_size_sample_vs_oracle(libszb):
    """BASELINE.json configs[3] at its model size (45 x 45 = 2025 machines, n = 20 250, fronts up to 454 in the
    per-CTA global workspace): a batch of bus-fault scenarios through the first 5 ms of the fault of scenario 0
    (Rodas rodas4, rtol 1e-3, atol 1e-6, hinit 1e-5); two scenarios against the oracle's Rodas loop
    (oracle/rodas.py) driven by the REFERENCE's own F / J of the same model (baseline/_ref, made_numerical):
    identical accepted / rejected step counts, trajectories 1e-8.  Size-independent properties on the rest: the
    unfaulted scenario stays at its equilibrium, a repeated scenario is bitwise equal."""
    import sys
    ref_dir = Path(__file__).resolve().parents[1] / "baseline" / "_ref"
    if not (ref_dir / "Solverz" / "__init__.py").exists():
        pytest.skip("baseline/_ref (private copy of the reference) is not present")
    sys.path.insert(0, str(ref_dir))
    import Solverz as SZ
    from oracle.rodas import rodas as oracle_rodas
    from solverz_b200 import Opt, Rodas, made_numerical
    from solverz_b200.solvers.rodas_param import Rodas_param
    from solverz_b200.synthetic import swing_dae_model, swing_fault_scenarios
    nx = 45
    m, info = swing_dae_model(nx, nx)
    dae, y0 = m.create_instance()
    nd = made_numerical(dae, y0, sparse=True)
    B = 16
    fmask, tst, tsv = swing_fault_scenarios(info["nb"], B)
    fmask[3] = 0.0                                       # scenario 3: no fault at all
    fmask[5], tst[5], tsv[5] = fmask[0], tst[0], tsv[0]  # scenario 5 repeats scenario 0
    nd.p["fmask"], nd.p["gf__ts_t"], nd.p["gf__ts_v"] = fmask, tst, tsv
    tsp = np.array([0.0, 0.05, 0.1, 0.1005, 0.102, 0.105])
    opt = Opt(rtol=1e-3, atol=1e-6, hinit=1e-5)
    sol = Rodas(nd, tsp, np.tile(y0.array, (B, 1)), opt)
    assert sol.stats.succeed.all()
    assert np.abs(sol.Y[3] - y0.array[None, :]).max() < 1e-10
    assert np.array_equal(sol.Y[5], sol.Y[0]) and np.array_equal(sol.stats.nstep[5:6], sol.stats.nstep[0:1])
    assert np.abs(sol.Y[0, -1] - y0.array).max() > 1e-3   # the fault moved the state
    # the reference's numerical functions of the same model, built with the reference's DSL
    mr, _ = swing_dae_model(nx, nx, ns=SZ)
    dr, y0r = mr.create_instance()
    ref = SZ.made_numerical(dr, y0r, sparse=True)
    np.testing.assert_allclose(y0r.array, y0.array, rtol=0, atol=1e-14)
    M = sps.csc_matrix(ref.M)
    rp = Rodas_param("rodas4")
    for b in (0, 1):
        ref.p["fmask"] = fmask[b].copy()                 # plain parameters are ndarrays in the reference's p dict
        gf = ref.p["gf"]                                 # the TimeSeriesParam object itself (get_v_t per call)
        gf.time_series[:] = tst[b]
        gf.v_series[:] = tsv[b]
        from scipy.interpolate import interp1d
        gf.vt = interp1d(gf.time_series, gf.v_series, kind='linear')
        gf.tend = gf.time_series[-1]
        F = lambda t, y: np.asarray(ref.F(t, y, ref.p)).copy()
        J = lambda t, y: sps.csc_matrix(ref.J(t, y, ref.p))
        r = oracle_rodas(F, J, M, tsp, y0.array, rp, rtol=1e-3, atol=1e-6, hinit=1e-5, dfdt_mode=0)
        assert int(sol.stats.nstep[b]) == r["stats"]["nstep"], (b, int(sol.stats.nstep[b]), r["stats"]["nstep"])
        assert int(sol.stats.nreject[b]) == r["stats"]["nreject"]
        d = np.abs(sol.Y[b] - r["Y"]).max(axis=1)
        assert d[:3].max() < 1e-8 and d.max() < 1e-5 * max(1.0, np.abs(r["Y"]).max()), (b, d)
    nd.close()


def test_rodas3_dense_output_ntrp2(libszb):
    """Dense output of rodas3 -- the cubic Hermite form ntrp2 (rodas.py:389-393) -- on the device, against the
    reference's own runs (tests/golden/rodas3_dense_ref.npz): 'fd' mode = the inline made_numerical path (two distinct
    F evaluations), 'module' mode = the rendered-module path, where both F calls of ntrp return one shared buffer
    (F0 is F1 there; the two fixtures differ by 1.7e-3, so the test tells the modes apart)."""
    from solverz_b200 import Opt, Rodas
    g = np.load(GOLD / "rodas3_dense_ref.npz")
    nd, y0 = _dae_test_model()
    for mode, key in (("fd", "inline"), ("module", "module")):
        opt = Opt(hinit=0.1, scheme="rodas3")
        opt.dfdt = mode
        sol = Rodas(nd, g["tspan"], y0, opt)
        got = [sol.stats.nstep, sol.stats.nfeval, sol.stats.nJeval, sol.stats.ndecomp, sol.stats.nreject, sol.stats.nsolve]
        assert got == g[f"stats_{key}"].tolist()
        assert np.max(np.abs(sol.Y.array - g[f"Y_{key}"])) < 1e-8, mode
    # a batch takes the same path
    sol = Rodas(nd, g["tspan"], np.tile(y0.array, (3, 1)), Opt(hinit=0.1, scheme="rodas3"))
    assert np.max(np.abs(sol.Y[1] - g["Y_module"])) < 1e-8
