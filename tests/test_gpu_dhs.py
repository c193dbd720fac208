"""GPU test of BASELINE.json configs[4] in miniature: the district-heating hydraulic-thermal FDAE
(``solverz_b200.synthetic.dhs_fdae_model``) as ONE batch of time-stepped load profiles through
``fdae_solver``, against the reference's fdae_solver run scenario by scenario on the same model written
with the reference's DSL (tests/golden/make_golden.py::golden_dhs)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def test_dhs_fdae_batch_matches_reference(libszb):
    from solverz_b200 import Opt, fdae_solver, made_numerical
    from solverz_b200.synthetic import dhs_fdae_model, dhs_load_profiles
    g = np.load(GOLD / "dhs_ref_3x3.npz")
    B = g["Y"].shape[0]
    m, info = dhs_fdae_model(3, 3)
    fdae, y0 = m.create_instance()
    mdl = made_numerical(fdae, y0, sparse=True)
    assert mdl.nstep == 1
    tst, tsv = dhs_load_profiles(B, info["dt"])
    mdl.p["sc__ts_t"], mdl.p["sc__ts_v"] = tst, tsv
    nt = g["Y"].shape[1] - 1
    sol = fdae_solver(mdl, [0, nt * info["dt"]], np.tile(y0.array, (B, 1)), Opt(step_size=info["dt"], ite_tol=1e-8))
    np.testing.assert_allclose(sol.T, g["T"], rtol=0, atol=1e-9)
    assert sol.stats.nstep == nt
    for b in range(B):
        ref = g["Y"][b]
        assert np.max(np.abs(sol.Y[b] - ref) / np.maximum(1.0, np.abs(ref))) < 1e-10, b
        st = g["stats"][b]
        assert [int(sol.stats.nfeval[b]), int(sol.stats.nJeval[b]), int(sol.stats.ndecomp[b]),
                int(sol.stats.nsolve[b])] == [st[1], st[2], st[3], st[5]], b
    assert sol.stats.succeed.all()
    # the loads really moved the network: flows and temperatures at the end differ from the start
    assert np.max(np.abs(sol.Y[:, -1] - sol.Y[:, 0])) > 1e-2


def test_dhs_fdae_medium_size_vs_reference_run(libszb):
    """configs[4] beyond the miniature: a 14 x 14 lattice (364 pipes, n = 1288; pipes re-oriented along their
    nominal flow) -- the batched fdae_solver against the REFERENCE's own fdae_solver run on the same model built with
    the reference's DSL (baseline/_ref, one scenario at a time, the load profile mutated in mdl.p): trajectories
    1e-9 relative, identical Newton counters per scenario."""
    import sys
    ref_dir = Path(__file__).resolve().parents[1] / "baseline" / "_ref"
    if not (ref_dir / "Solverz" / "__init__.py").exists():
        pytest.skip("baseline/_ref (private copy of the reference) is not present")
    sys.path.insert(0, str(ref_dir))
    import Solverz as SZ
    from solverz_b200 import Opt, fdae_solver, made_numerical
    from solverz_b200.synthetic import dhs_fdae_model, dhs_load_profiles
    nx, B, nt = 14, 5, 8
    m, info = dhs_fdae_model(nx, nx)
    fdae, y0 = m.create_instance()
    mdl = made_numerical(fdae, y0, sparse=True)
    dt = info["dt"]
    tst, tsv = dhs_load_profiles(B, dt)
    mdl.p["sc__ts_t"], mdl.p["sc__ts_v"] = tst, tsv
    sol = fdae_solver(mdl, [0, nt * dt], np.tile(y0.array, (B, 1)), Opt(step_size=dt, ite_tol=1e-8))
    assert sol.stats.succeed.all()
    mr, _ = dhs_fdae_model(nx, nx, ns=SZ)
    fr, y0r = mr.create_instance()
    ref = SZ.made_numerical(fr, y0r, sparse=True)
    np.testing.assert_allclose(y0r.array, y0.array, rtol=0, atol=1e-12)
    for b in range(B):
        ref.p["sc"] = SZ.TimeSeriesParam('sc', v_series=tsv[b], time_series=tst[b])
        rs = SZ.fdae_solver(ref, [0, nt * dt], y0r, SZ.Opt(step_size=dt, ite_tol=1e-8))
        Yr = rs.Y.array
        assert np.max(np.abs(sol.Y[b] - Yr) / np.maximum(1.0, np.abs(Yr))) < 1e-9, b
        assert int(sol.stats.nsolve[b]) == int(rs.stats.nsolve) and int(sol.stats.nfeval[b]) == int(rs.stats.nfeval), b
