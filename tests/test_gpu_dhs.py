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
